"""Drop-in for the reference's ``src/model/losses.py`` on CUDA devices.

Same public names, arguments and return values:

  YoloDFLQFLoss(num_classes=171, lambda_box=1.5, lambda_cls=1.0, lambda_dfl=1.5, reg_max=16)
      .forward(preds, gt_boxes_list, anchors, strides) -> (loss, {"total_loss","box_loss","cls_loss"})
  bbox_iou(box1, box2)                                  (reference :9-40)
  quality_focal_loss(pred_scores, target_scores, beta)  (reference :46-57)
  distribution_focal_loss(pred_dist, target_val)        (reference :63-78)

``forward`` runs the whole of reference lines 140-281 *and* the backward of it in three CUDA
launches (``csrc/loss.cu``): the gradient w.r.t. ``preds`` is produced during the forward call and
handed to autograd by ``_FusedLoss.backward``.  The reference's behaviours that decide results are
kept — see SURVEY.md §0.2 (Q1-Q7, Q16) and DESIGN.md.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _cabi

__all__ = ["YoloDFLQFLoss", "bbox_iou", "quality_focal_loss", "distribution_focal_loss", "pack_gt", "pack_gt_host",
           "PackedGT", "fused_loss", "fused_tal_loss"]


# ----------------------------------------------------------------------------------------------
# host logic: GT list -> one concatenated tensor + offsets (no per-image device work, no sync)
# ----------------------------------------------------------------------------------------------
_pack_rings = {}        # (device index, words) -> {"bufs": [pinned int32 staging buffers], "events": [...], "next": i}


def _pack_gt_gather(gt_boxes_list, device):
    """Fast path of ``pack_gt``: every entry is an fp32 ``(Mi, >=5)`` tensor on ``device`` with unit column stride (what the
    reference's loop builds, train_model.py:236).  The host writes one table -- offsets and, per image, (pointer, row pitch,
    first row, rows) -- into a pinned staging buffer, copies it with one asynchronous H2D copy, and ``yb_gather_gt`` gathers
    the rows in one launch: no per-image slice / cast / cat on the host.  Returns ``None`` when an entry does not qualify."""
    n = len(gt_boxes_list)
    words = (n + 1) + (n + 1) % 2 + 6 * n                 # offsets (padded to 8 bytes) + 24-byte table entries, in int32 words
    ptrs, pitch, first, count = [], [], [], []
    total = 0
    f32, idx = torch.float32, device.index
    for g in gt_boxes_list:                                # (kept lean: this loop IS the host cost of the list interface)
        if not isinstance(g, torch.Tensor) or g.dtype is not f32 or not g.is_cuda or g.get_device() != idx:
            return None
        sh = g.shape
        m = sh[0] if len(sh) == 2 else -1
        if m <= 0 or sh[1] == 0:
            if g.numel() != 0:
                return None
            ptrs.append(0); pitch.append(5); first.append(total); count.append(0)
            continue
        st = g.stride()
        if sh[1] < 5 or st[1] != 1:
            return None
        ptrs.append(g.data_ptr()); pitch.append(st[0]); first.append(total); count.append(m)
        total += m
    key = (device.index, words)
    ring = _pack_rings.get(key)
    if ring is None:
        ring = _pack_rings[key] = {"bufs": [torch.empty(words, dtype=torch.int32).pin_memory() for _ in range(4)],
                                   "events": [None] * 4, "next": 0}
    i = ring["next"]
    ring["next"] = (i + 1) % 4
    if ring["events"][i] is not None:
        ring["events"][i].synchronize()                     # the copy that last used this staging buffer (long done)
    stage = ring["bufs"][i]
    host = stage.numpy()
    t0 = (n + 1) + (n + 1) % 2
    host[:n] = first
    host[n] = total
    table = host[t0:].reshape(n, 6)
    table[:, :2].view("uint64")[:, 0] = ptrs
    table[:, 2] = pitch
    table[:, 3] = first
    table[:, 4] = count
    table[:, 5] = 0
    dev_buf = stage.to(device, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    ring["events"][i] = ev
    gt = torch.empty(total, 5, dtype=torch.float32, device=device)
    if total:
        with torch.cuda.device(device):
            rc = _cabi.lib().yb_gather_gt(dev_buf.data_ptr() + 4 * t0, n, _cabi.ptr(gt), _cabi.stream_ptr(device))
        _cabi.check(rc, "yb_gather_gt")
    return gt, dev_buf[: n + 1], count


def pack_gt(gt_boxes_list: Sequence[torch.Tensor], device) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """``list[(Mi, 5)]`` -> ``(gt (sum Mi, 5) fp32, offsets (N+1,) int32, counts)`` on ``device``.

    The counts come from tensor *shapes*, so nothing here waits for the GPU.  Boxes are cast to
    fp32 as the reference does (``gt_boxes[:, 0:4].to(preds.dtype)`` with preds already float,
    losses.py:208); the class column is truncated toward zero inside the kernel (``.long()``, :257).
    fp32 tensors already on ``device`` -- the reference's own calling convention -- are gathered by one kernel launch
    (``_pack_gt_gather``); anything else (CPU tensors, other dtypes, odd strides) goes through torch ops.
    """
    device = torch.device(device)
    if device.type == "cuda" and len(gt_boxes_list) > 0:
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        fast = _pack_gt_gather(gt_boxes_list, device)
        if fast is not None:
            return fast
    counts = []
    parts = []
    for i, g in enumerate(gt_boxes_list):
        if not isinstance(g, torch.Tensor):
            raise TypeError(f"gt_boxes_list[{i}] must be a tensor, got {type(g).__name__}")
        if g.numel() == 0:
            counts.append(0)
            continue
        if g.dim() != 2 or g.shape[1] < 5:
            raise ValueError(f"gt_boxes_list[{i}] must have shape (Mi, 5), got {tuple(g.shape)}")
        counts.append(int(g.shape[0]))
        parts.append(g[:, :5].to(device=device, dtype=torch.float32))
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    if parts:
        gt = torch.cat(parts, 0).contiguous() if len(parts) > 1 else parts[0].contiguous()
    else:
        gt = torch.zeros(0, 5, dtype=torch.float32, device=device)
    off = torch.tensor(offsets, dtype=torch.int32).to(device, non_blocking=True)
    return gt, off, counts


class PackedGT:
    """The GT wire format of the kernels: ``gt (sum Mi, 5)`` fp32, ``offsets (N+1,)`` int32, per-image counts.
    ``YoloDFLQFLoss.forward`` accepts it in place of the per-image list (one pinned
    staging buffer and ONE host-to-device copy per batch instead of N small ones)."""

    __slots__ = ("gt", "offsets", "counts")

    def __init__(self, gt, offsets, counts):
        self.gt, self.offsets, self.counts = gt, offsets, list(counts)

    def __len__(self):
        return len(self.counts)

    def to(self, device, non_blocking=True):
        return PackedGT(self.gt.to(device, non_blocking=non_blocking), self.offsets.to(device, non_blocking=non_blocking),
                        self.counts)


def pack_gt_host(gt_boxes_list: Sequence[torch.Tensor], pin_memory: bool = True) -> PackedGT:
    """Collate-side packing: CPU ``list[(Mi, 5)]`` -> one (pinned) ``(sum Mi, 5)`` buffer + offsets.
    Use in a ``collate_fn`` and move the result with ``.to(device)``: two async copies per batch,
    independent of the batch size (the reference does N copies, src/training/train_model.py:236)."""
    counts = [0 if (g is None or g.numel() == 0) else int(g.shape[0]) for g in gt_boxes_list]
    total = sum(counts)
    gt = torch.empty(total, 5, dtype=torch.float32)
    off = torch.zeros(len(counts) + 1, dtype=torch.int32)
    pos = 0
    for i, (g, c) in enumerate(zip(gt_boxes_list, counts)):
        if c:
            if g.dim() != 2 or g.shape[1] < 5:
                raise ValueError(f"gt_boxes_list[{i}] must have shape (Mi, 5), got {tuple(g.shape)}")
            gt[pos:pos + c] = g[:, :5].to(torch.float32)
        pos += c
        off[i + 1] = pos
    if pin_memory and torch.cuda.is_available():
        gt, off = gt.pin_memory(), off.pin_memory()
    return PackedGT(gt, off, counts)


def _workspace(n_bytes: int, device) -> torch.Tensor:
    return torch.empty(max(n_bytes, 16), dtype=torch.uint8, device=device)


_ws_cache = {}


def _stream_workspace(n_bytes: int, device) -> torch.Tensor:
    """Scratch buffer reused by consecutive calls on one (device, stream): calls on a stream are
    ordered, so the next call may overwrite the previous call's scratch."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < n_bytes:
        buf = torch.empty(max(n_bytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


_clean_ws_cache = {}


def _clean_workspace(n_bytes: int, device, tag: str) -> torch.Tensor:
    """A scratch buffer per (device, stream, entry point) that the kernels keep ZEROED between calls: allocated as zeros
    (on the current stream, so ordered before its first use) and handed only to an entry point that wipes what it used
    before its last kernel ends (``YB_LOSS_WS_CLEAN`` / ``YB_TAL_WS_CLEAN``) -- the step then needs no memset node."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _clean_ws_cache.get(key)
    if buf is None or buf.numel() < n_bytes:
        buf = torch.zeros(max(n_bytes, 1 << 16), dtype=torch.uint8, device=device)
        _clean_ws_cache[key] = buf
    return buf


class _keeps_workspace_clean:
    """``with _keeps_workspace_clean(device, tag): ...`` around the calls that use a clean workspace: if anything raises in
    between (an ABI error, a failing collective between ``yb_tal_assign`` and ``yb_tal_loss``), what the kernels left in the
    buffer is unknown -- forget it, the next call zeroes a fresh one."""

    def __init__(self, device, tag: str):
        self.key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is not None:
            _clean_ws_cache.pop(self.key, None)
        return False


def _as_offsets(gt: torch.Tensor, gt_offsets: torch.Tensor, n: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """The GT wire format as the kernels read it: ``gt (sum Mi, 5)`` fp32 and ``offsets (N+1,)`` int32, contiguous, on
    ``device``.  A CPU (pinned) ``PackedGT`` straight from ``collate_fn_packed`` is moved, not dereferenced as a device
    pointer.  Shape mistakes raise here; the VALUES of device-resident offsets cannot be checked without a sync."""
    if gt_offsets.dim() != 1 or gt_offsets.numel() != n + 1:
        raise ValueError(f"gt_offsets must have shape ({n + 1},) for a batch of {n}, got {tuple(gt_offsets.shape)}")
    if gt.dim() != 2 or (gt.shape[0] and gt.shape[1] != 5):
        raise ValueError(f"gt must have shape (sum Mi, 5), got {tuple(gt.shape)}")
    if not gt_offsets.is_cuda:                                  # host offsets are free to validate
        off = gt_offsets.tolist()
        if off[0] != 0 or off[-1] != gt.shape[0] or any(b < a for a, b in zip(off, off[1:])):
            raise ValueError(f"gt_offsets must rise from 0 to {gt.shape[0]} (the number of gt rows), got {off[:4]}...{off[-1]}")
    if gt_offsets.device != device or gt_offsets.dtype != torch.int32 or not gt_offsets.is_contiguous():
        gt_offsets = gt_offsets.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    if gt.device != device or gt.dtype != torch.float32 or not gt.is_contiguous():
        gt = gt.detach().to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    return gt, gt_offsets


def _as_f32(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype == torch.float32 and t.device == device and t.is_contiguous() and not t.requires_grad:
        return t
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def fused_loss(preds: torch.Tensor, gt: torch.Tensor, gt_offsets: torch.Tensor, gmax: int, anchors: torch.Tensor,
               strides: torch.Tensor, num_classes: int, lambda_cls: float, lambda_dfl: float, reg_max: int = 16,
               want_grad: bool = True, want_trace: bool = False, flags: int = 0, stage_events=None, grid_hint="auto"):
    """One call of ``yb_loss_fwd_bwd``.  Returns ``(out_loss (8,), grad or None, trace dict)``.

    ``out_loss`` = [total, mean DFL, mean QFL, #matched anchors, 0, 0, in-kernel dependency timed out (never, see
    csrc/common.cuh), #GT rows with a class id outside [0, nc)] on the device.  ``trace`` (``want_trace``) holds the per-GT matched anchor / IoU and per-image
    loss terms the parity tests compare against the oracle.  ``flags``: ``_cabi.YB_LOSS_NO_PRUNE`` /
    ``YB_LOSS_SPLIT_LAUNCH`` (test / profiling aids).  ``stage_events``: three ``torch.cuda.Event(enable_timing=True)``
    recorded around the launch(es) (bench.py's roofline leg).  ``grid_hint``: ``"auto"`` (describe the anchors as a
    pyramid of grids once per size: the launch then bounds every GT's nearest-centre distance from a few probe anchors and
    prunes from its first tile on), ``None`` or a ``_cabi.TalGrid``; results never depend on it.
    """
    _cabi.require_cuda(preds, "preds")
    if preds.dim() != 3:
        raise ValueError(f"preds must be (N, 4*reg_max + nc, A), got {tuple(preds.shape)}")
    n, c, a = preds.shape
    if c != 4 * reg_max + num_classes:
        raise ValueError(f"preds has {c} channels, expected 4*{reg_max} + {num_classes}")
    dev = preds.device
    dt = _cabi.dtype_code(preds.dtype)
    x = preds.detach() if preds.requires_grad else preds
    if not x.is_contiguous():
        x = x.contiguous()
    anc = _as_f32(anchors, dev)
    st = _as_f32(strides, dev)
    if anc.shape != (2, a) or st.numel() != a:
        raise ValueError(f"anchors must be (2, {a}) and strides (1, {a}); got {tuple(anc.shape)}, {tuple(st.shape)}")
    gt, gt_offsets = _as_offsets(gt, gt_offsets, n, dev)
    gt_total = int(gt.shape[0])
    lib = _cabi.lib()
    if int(flags) & _cabi.YB_LOSS_SPLIT_LAUNCH:                 # the profiling split leaves its counters behind
        ws = _stream_workspace(lib.yb_loss_workspace_bytes(n, a, gt_total, dt), dev)
    else:
        ws = _clean_workspace(lib.yb_loss_workspace_bytes(n, a, gt_total, dt), dev, "loss")
        flags = int(flags) | _cabi.YB_LOSS_WS_CLEAN
    ev = None
    if stage_events is not None:
        import ctypes
        for e in stage_events:                      # torch creates the CUDA event lazily, on the first record
            if not e.cuda_event:
                e.record()
        ev = (ctypes.c_void_p * 3)(*[e.cuda_event for e in stage_events])
    grad = torch.empty_like(x) if want_grad else None
    out = torch.empty(8, dtype=torch.float32, device=dev)
    trace = {}
    idx = iou = per_image = None
    if want_trace:
        idx = torch.empty(max(gt_total, 1), dtype=torch.int32, device=dev)
        iou = torch.empty(max(gt_total, 1), dtype=torch.float32, device=dev)
        per_image = torch.empty(2, n, dtype=torch.float32, device=dev)
        trace = {"idx": idx[:gt_total], "iou": iou[:gt_total], "dfl_per_image": per_image[0], "cls_per_image": per_image[1]}
    hint = _grid_hint_for(anc, st, exact=False) if isinstance(grid_hint, str) else grid_hint
    with torch.cuda.device(dev), _keeps_workspace_clean(dev, "loss"):
        import ctypes
        rc = lib.yb_loss_fwd_bwd(_cabi.ptr(x), dt, n, num_classes, reg_max, a, _cabi.ptr(anc), _cabi.ptr(st),
                                 _cabi.ptr(gt) if gt_total else None, _cabi.ptr(gt_offsets), gt_total, gmax,
                                 float(lambda_cls), float(lambda_dfl), _cabi.ptr(grad), _cabi.ptr(out),
                                 _cabi.ptr(idx), _cabi.ptr(iou), _cabi.ptr(per_image), _cabi.ptr(ws), ws.numel(),
                                 int(flags), ctypes.byref(hint) if hint is not None else None, ev, _cabi.stream_ptr(dev))
        _cabi.check(rc, "yb_loss_fwd_bwd")
    return out, grad, trace


def _take_grad(ctx):
    """The gradient produced during the forward call is handed over once (it is scaled in place).  A second
    backward through the same graph raises, as autograd does for freed buffers, instead of returning None."""
    if getattr(ctx, "grad_taken", False):
        raise RuntimeError("Trying to backward through the fused loss a second time: the gradient buffer was handed "
                           "over (and scaled in place) by the first backward; run forward again")
    ctx.grad_taken = True
    g = ctx.grad
    ctx.grad = None
    return g


def _raise_on_bad_class(n_bad: float, num_classes: int):
    # the reference's scatter_ raises on these (src/model/losses.py:260); the kernels clamp and count
    if n_bad:
        raise RuntimeError(f"index out of range: {int(n_bad)} ground-truth row(s) carry a class id outside [0, {num_classes})")


def _raise_on_stall(flag: float, who: str):
    # a consumer CTA of the fused launch gave up waiting for its producers (csrc/common.cuh::dep_wait): the block
    # dispatch order the launch relies on did not hold -- fail loudly, the loss of that call is NaN
    if flag:
        _clean_ws_cache.clear()                     # what the stalled launch left in its workspace is unknown
        raise RuntimeError(f"{who}: an in-kernel dependency timed out (blocks were not dispatched in index order)")


class LossDict(dict):
    """The ``{"total_loss", "box_loss", "cls_loss"[, "dfl_loss"]}`` dict of Python floats that ``forward`` returns
    (src/model/losses.py:277-281), fetched LAZILY: ``forward`` only queues one asynchronous device-to-host copy of the
    8-float statistics vector into pinned memory, and the first read of a value waits for it.  The reference's training
    loop reads the dict after ``loss.backward()`` and ``optimizer.step()`` (train_model.py:247-258), so the host no longer
    stalls inside ``forward`` for the kernels to finish (the reference's three ``.item()`` calls do).  What the kernels
    report about the inputs -- class ids outside ``[0, nc)``, on which the reference's ``scatter_`` raises -- is raised by
    the same first read.  Every way of reading values goes through ``_fetch``; keys, ``len`` and ``in`` need no data."""

    _ring = {}                          # device index -> [pinned (8,) buffers], recycled round-robin
    _owner = {}                         # id(buffer) -> weakref to the LossDict still waiting on it

    def __init__(self, stats: torch.Tensor, names, check):
        super().__init__((k, None) for k in names)
        self._names = tuple(names)
        self._check = check
        ring = LossDict._ring.setdefault(stats.device.index, {"bufs": [torch.empty(8, dtype=torch.float32).pin_memory()
                                                                       for _ in range(8)], "next": 0})
        buf = ring["bufs"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % len(ring["bufs"])
        prev = LossDict._owner.get(id(buf))
        prev = prev() if prev is not None else None
        if prev is not None:
            prev._fetch()                                   # an older dict nobody has read yet still owns this buffer
        import weakref
        LossDict._owner[id(buf)] = weakref.ref(self)
        self._buf = buf
        buf.copy_(stats, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(stats.device))

    def _fetch(self):
        if self._buf is None:
            return
        self._event.synchronize()
        host = self._buf.tolist()
        self._buf = self._event = None
        for i, k in enumerate(self._names):
            dict.__setitem__(self, k, host[i])
        check, self._check = self._check, None
        check(host)                                         # may raise (the values are in place either way)

    def __getitem__(self, k):
        self._fetch()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        self._fetch()
        return dict.get(self, k, default)

    def __iter__(self):                 # (also keeps dict(d) / {**d} off the raw-storage fast path)
        return iter(self._names)

    def keys(self):
        return dict.keys(self)

    def values(self):
        self._fetch()
        return dict.values(self)

    def items(self):
        self._fetch()
        return dict.items(self)

    def copy(self):
        self._fetch()
        return dict(dict.items(self))

    def __eq__(self, other):
        self._fetch()
        if isinstance(other, LossDict):
            other._fetch()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._fetch()
        return dict.__repr__(self)

    def __reduce__(self):
        self._fetch()
        return (dict, (dict(dict.items(self)),))


class _FusedLoss(torch.autograd.Function):
    """Autograd bridge: forward launches the fused fwd+bwd kernels, backward hands the stored
    gradient over (scaled in place by ``grad_output`` only when that is not exactly 1)."""

    @staticmethod
    def forward(ctx, preds, gt, gt_offsets, gmax, anchors, strides, num_classes, lambda_cls, lambda_dfl, reg_max, need,
                holder):
        out, grad, _ = fused_loss(preds, gt, gt_offsets, gmax, anchors, strides, num_classes, lambda_cls, lambda_dfl,
                                  reg_max, want_grad=need)
        ctx.grad = grad
        holder.append(out)              # the 8-float stats vector; the loss itself is its element 0
        return out[0]

    @staticmethod
    def backward(ctx, grad_total):
        g = _take_grad(ctx)
        if g is None:
            return (None,) * 12
        scale = grad_total.detach().to(device=g.device, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(g.device):
            rc = _cabi.lib().yb_scale_grad(_cabi.ptr(g), _cabi.dtype_code(g.dtype), g.numel(), _cabi.ptr(scale),
                                           _cabi.stream_ptr(g.device))
        _cabi.check(rc, "yb_scale_grad")
        return (g,) + (None,) * 11


# ----------------------------------------------------------------------------------------------
# host logic: the anchors' grid structure, as a hint the kernels verify (never a precondition)
# ----------------------------------------------------------------------------------------------
_grid_hints = {}        # (device index, A) -> TalGrid | None (the anchors are not a pyramid of grids)
_grid_strikes = {}


def build_grid_hint(anchors: torch.Tensor, strides: torch.Tensor, exact: bool = True):
    """Describe ``anchors (2, A)`` / ``strides (1, A)`` as the reference's pyramid of regular grids
    (``make_anchors``, model_utils.py:60-70: per level x fastest, ``(x0 + col, y0 + row)``, one stride per level), or
    return ``None`` when they are not one.  Pure host logic on a CPU copy; called once per anchor-set size, the
    device re-verifies the result on every call.  ``exact=False`` accepts coordinates within one cell of the grid
    (anchors built in bf16): good enough for ``yb_loss_fwd_bwd``'s probe role, which only
    needs a cell NEAR each box centre; the task-aligned path verifies its hint bit for bit and takes exact ones only."""
    anc = anchors.detach().float().cpu().reshape(2, -1)
    st = strides.detach().float().cpu().reshape(-1)
    a = st.numel()
    if a == 0 or anc.shape[1] != a:
        return None
    cuts = [0] + (torch.nonzero(st[1:] != st[:-1])[:, 0] + 1).tolist() + [a]
    if len(cuts) - 1 > _cabi.TAL_MAX_LEVELS:
        return None
    hint = _cabi.TalGrid()
    hint.n_levels = len(cuts) - 1
    for l, (lo, hi) in enumerate(zip(cuts, cuts[1:])):
        ax, ay = anc[0, lo:hi], anc[1, lo:hi]
        rows = torch.nonzero(ay[1:] != ay[:-1])[:, 0]
        w = int(rows[0]) + 1 if rows.numel() else hi - lo
        if (hi - lo) % w or not st[lo] > 0:
            return None
        h = (hi - lo) // w
        col = torch.arange(w, dtype=torch.float32).repeat(h)
        row = torch.arange(h, dtype=torch.float32).repeat_interleave(w)
        if exact:
            if not (torch.equal(ax, ax[0] + col) and torch.equal(ay, ay[0] + row)):
                return None
        elif not ((ax - (ax[0] + col)).abs().max() <= 1.0 and (ay - (ay[0] + row)).abs().max() <= 1.0):
            return None
        hint.start[l], hint.w[l], hint.h[l] = lo, w, h
        hint.stride[l], hint.x0[l], hint.y0[l] = float(st[lo]), float(ax[0]), float(ay[0])
    return hint


def _grid_hint_for(anc: torch.Tensor, st: torch.Tensor, exact: bool = True):
    key = (anc.device.index, anc.shape[1]) if exact else (anc.device.index, anc.shape[1], "approx")
    if key not in _grid_hints:
        _grid_hints[key] = build_grid_hint(anc, st, exact)    # one device-to-host copy, the first time this size is seen
    return _grid_hints[key]


def grid_hint_rejected(anc_device_index: int, a: int):
    """The kernels reported (out_loss[6]) that the cached hint does not describe the anchors they were given: forget it,
    so that the next call rebuilds it from the new anchors; give up on a size whose fresh hints keep failing."""
    key = (anc_device_index, a)
    _grid_strikes[key] = _grid_strikes.get(key, 0) + 1
    if _grid_strikes[key] >= 3:
        _grid_hints[key] = None
    else:
        _grid_hints.pop(key, None)


def fused_tal_loss(preds: torch.Tensor, gt: torch.Tensor, gt_offsets: torch.Tensor, anchors: torch.Tensor,
                   strides: torch.Tensor, num_classes: int, lambda_box: float, lambda_cls: float, lambda_dfl: float,
                   reg_max: int = 16, topk: int = 10, alpha: float = 0.5, beta: float = 6.0, want_grad: bool = True,
                   want_trace: bool = False, sync_normalizer: bool = True, cls_loss: str = "bce",
                   vfl_alpha: float = 0.75, vfl_gamma: float = 2.0, grid_hint="auto", exchange="default"):
    """Task-aligned variant (``yb_tal_assign`` + ``yb_tal_loss``).  Not in the reference: specified by
    ``oracle/tal_oracle.py`` (SURVEY.md §8(a')).

    Between the two calls the normaliser ``sum(target scores)`` is all-reduced (SUM / world) when a
    process group is initialised and ``sync_normalizer`` is set — the path's one real exchange step.
    ``cls_loss="vfl"`` weights the class term varifocally (``yb_tal_params.vfl``): background cells by
    ``vfl_alpha * sigmoid(x) ** vfl_gamma``, the positive cell of a foreground anchor by its target score.
    ``exchange``: a ``training.distributed_setup.PeerExchange`` fuses that step into the kernels (peer stores over
    NVLink + a poll, no collective call); ``"default"`` takes the one ``enable_peer_exchange()`` installed, if any.
    ``grid_hint``: ``"auto"`` (describe the anchors as a pyramid of grids once per size and let the kernels verify it on
    every call), ``None`` (structure-free candidate scan) or a ``_cabi.TalGrid``.  Results never depend on it.
    Returns ``(out_loss (8,) [total, box, cls, dfl, normaliser, #fg, hint rejected, #bad class ids], grad or None, trace)``.
    """
    _cabi.require_cuda(preds, "preds")
    if cls_loss not in ("bce", "vfl"):
        raise ValueError(f"cls_loss must be 'bce' or 'vfl', got {cls_loss!r}")
    n, c, a = preds.shape
    if c != 4 * reg_max + num_classes:
        raise ValueError(f"preds has {c} channels, expected 4*{reg_max} + {num_classes}")
    dev = preds.device
    dt = _cabi.dtype_code(preds.dtype)
    x = preds.detach() if preds.requires_grad else preds
    if not x.is_contiguous():
        x = x.contiguous()
    anc, st = _as_f32(anchors, dev), _as_f32(strides, dev)
    gt, gt_offsets = _as_offsets(gt, gt_offsets, n, dev)
    gt_total = int(gt.shape[0])
    lib = _cabi.lib()
    ws = _clean_workspace(lib.yb_tal_workspace_bytes(n, a, gt_total, dt, topk), dev, "tal")
    stats = torch.empty(8, dtype=torch.float32, device=dev)
    asg = tsc = None
    if want_trace:
        asg = torch.empty(n, a, dtype=torch.int32, device=dev)
        tsc = torch.empty(n, a, dtype=torch.float32, device=dev)
    gt_ptr = _cabi.ptr(gt) if gt_total else None
    import ctypes
    hint = _grid_hint_for(anc, st) if isinstance(grid_hint, str) else grid_hint
    if isinstance(exchange, str):
        from ..training.distributed_setup import default_peer_exchange
        exchange = default_peer_exchange()
    px = exchange.next_step() if (exchange is not None and sync_normalizer) else None
    px_ref = ctypes.byref(px) if px is not None else None
    params = _cabi.TalParams(int(topk), float(alpha), float(beta), float(lambda_box), float(lambda_cls), float(lambda_dfl),
                             int(cls_loss == "vfl"), float(vfl_alpha), float(vfl_gamma), _cabi.YB_TAL_WS_CLEAN)
    with _keeps_workspace_clean(dev, "tal"):       # an exception between the two calls leaves the counters armed
        with torch.cuda.device(dev):
            rc = lib.yb_tal_assign(_cabi.ptr(x), dt, n, num_classes, reg_max, a, _cabi.ptr(anc), _cabi.ptr(st), gt_ptr,
                                   _cabi.ptr(gt_offsets), gt_total, ctypes.byref(params),
                                   ctypes.byref(hint) if hint is not None else None, px_ref, _cabi.ptr(stats),
                                   _cabi.ptr(asg), _cabi.ptr(tsc), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(dev))
        _cabi.check(rc, "yb_tal_assign")
        tss = stats[:1]
        if px is None and sync_normalizer and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            # [sum of target scores, #foreground] -> mean over the ranks, in place: one collective, no extra kernels.
            # Nothing of the step is left to overlap it with: everything that does not need the normaliser has already run.
            if torch.distributed.get_backend() == "nccl":
                torch.distributed.all_reduce(stats[:2], op=torch.distributed.ReduceOp.AVG)
            else:                                                   # gloo has no AVG
                torch.distributed.all_reduce(stats[:2])
                stats[:2] /= torch.distributed.get_world_size()
            tss = stats[:1]
        grad = torch.empty_like(x) if want_grad else None
        out = torch.empty(8, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.yb_tal_loss(_cabi.ptr(x), dt, n, num_classes, reg_max, a, gt_total, ctypes.byref(params), _cabi.ptr(tss),
                                 px_ref, _cabi.ptr(grad), _cabi.ptr(out), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(dev))
        _cabi.check(rc, "yb_tal_loss")
    trace = {"assigned_gt": asg, "target_score": tsc, "stats": stats} if want_trace else {}
    return out, grad, trace


class _FusedTalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, gt, gt_offsets, anchors, strides, num_classes, lambdas, reg_max, tal, need, holder):
        out, grad, _ = fused_tal_loss(preds, gt, gt_offsets, anchors, strides, num_classes, lambdas[0], lambdas[1],
                                      lambdas[2], reg_max, tal["topk"], tal["alpha"], tal["beta"], want_grad=need,
                                      sync_normalizer=tal["sync_normalizer"], cls_loss=tal.get("cls_loss", "bce"),
                                      vfl_alpha=tal.get("vfl_alpha", 0.75), vfl_gamma=tal.get("vfl_gamma", 2.0))
        ctx.grad = grad
        holder.append(out)
        return out[0]

    @staticmethod
    def backward(ctx, grad_total):
        g = _take_grad(ctx)
        if g is None:
            return (None,) * 11
        scale = grad_total.detach().to(device=g.device, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(g.device):
            rc = _cabi.lib().yb_scale_grad(_cabi.ptr(g), _cabi.dtype_code(g.dtype), g.numel(), _cabi.ptr(scale),
                                           _cabi.stream_ptr(g.device))
        _cabi.check(rc, "yb_scale_grad")
        return (g,) + (None,) * 10


class YoloDFLQFLoss(nn.Module):
    """Same constructor and ``forward`` contract as the reference class (losses.py:84-281).

    ``lambda_box`` is accepted and, as in the reference, never used (losses.py:88, :275).
    ``forward`` returns ``(total_loss, {"total_loss", "box_loss", "cls_loss"})`` with the dict holding
    Python floats; they are fetched with ONE asynchronous device-to-host copy that the first read of a value waits for
    (``LossDict``), instead of three ``.item()`` syncs inside ``forward``.
    ``last_stats`` keeps the 8-float device vector [total, dfl, cls, #matched, ...] of the last call
    for ``training.distributed_setup.reduce_loss_stats``.

    ``assigner="tal"`` (not in the reference; BASELINE.json's north_star) switches to the task-aligned
    assigner with CIoU + DFL + BCE losses (``topk``, ``alpha``, ``beta``; ``cls_loss="vfl"`` weights the
    class term varifocally with ``vfl_alpha``, ``vfl_gamma``); the dict then also carries
    ``"dfl_loss"`` and ``"box_loss"`` is the CIoU term.  Under DDP the normaliser is all-reduced.
    """

    def __init__(self, num_classes=171, lambda_box=1.5, lambda_cls=1.0, lambda_dfl=1.5, reg_max=16,
                 assigner="nearest_center", topk=10, alpha=0.5, beta=6.0, sync_normalizer=True, cls_loss="bce",
                 vfl_alpha=0.75, vfl_gamma=2.0):
        super().__init__()
        if assigner not in ("nearest_center", "tal"):
            raise ValueError(f"assigner must be 'nearest_center' (the reference's behaviour) or 'tal', got {assigner!r}")
        self.num_classes = num_classes
        self.lambda_box = lambda_box
        self.lambda_cls = lambda_cls
        self.lambda_dfl = lambda_dfl
        if reg_max != 16:
            # the reference's Head is hard-wired to 16 bins (src/model/head.py:35) and the kernels are built for that
            raise ValueError(f"reg_max must be 16 (the head's DFL channel count, src/model/head.py:35), got {reg_max}")
        self.reg_max = reg_max
        self.assigner = assigner
        if cls_loss not in ("bce", "vfl"):
            raise ValueError(f"cls_loss must be 'bce' or 'vfl' (task-aligned variant only), got {cls_loss!r}")
        self.tal = {"topk": topk, "alpha": alpha, "beta": beta, "sync_normalizer": sync_normalizer,
                    "cls_loss": cls_loss, "vfl_alpha": vfl_alpha, "vfl_gamma": vfl_gamma}
        self.last_stats = None

    def forward(self, preds, gt_boxes_list, anchors, strides):
        n = preds.shape[0]
        if n == 0:                                  # the reference returns a zero loss and an empty dict (losses.py:268-269)
            return torch.tensor(0.0, device=preds.device), {}
        if preds.dtype not in (torch.float32, torch.bfloat16):
            # fp16 (autocast + GradScaler) / fp64 head outputs: the kernels read fp32 or bf16, so do what the
            # reference does (losses.py:142 `.float()`); autograd casts the gradient back
            preds = preds.float()
        if len(gt_boxes_list) != n:
            raise IndexError(f"gt_boxes_list has {len(gt_boxes_list)} entries for a batch of {n}")
        if isinstance(gt_boxes_list, PackedGT):
            packed = gt_boxes_list if gt_boxes_list.gt.device == preds.device else gt_boxes_list.to(preds.device)
            gt, off, counts = packed.gt, packed.offsets, packed.counts
        else:
            gt, off, counts = pack_gt(gt_boxes_list, preds.device)
        if self.assigner == "tal":
            # not in the reference: task-aligned assigner + CIoU / DFL / BCE, here lambda_box IS used
            need_grad = torch.is_grad_enabled() and preds.requires_grad
            holder = []
            total = _FusedTalLoss.apply(preds, gt, off, anchors, strides, self.num_classes,
                                        (self.lambda_box, self.lambda_cls, self.lambda_dfl), self.reg_max, self.tal,
                                        need_grad, holder)
            stats = self.last_stats = holder[0]
            nc, dev_index, a = self.num_classes, preds.device.index, preds.shape[2]

            def check_tal(host):
                if host[6]:                         # the anchors changed under a cached grid hint (the result is still right)
                    grid_hint_rejected(dev_index, a)
                _raise_on_bad_class(host[7], nc)
            return total, LossDict(stats, ("total_loss", "box_loss", "cls_loss", "dfl_loss"), check_tal)
        if n > 0 and sum(counts) == 0:
            # the reference fails here: total_dfl is still the python float 0.0 (losses.py:271-279)
            raise AttributeError("'float' object has no attribute 'detach'")
        need_grad = torch.is_grad_enabled() and preds.requires_grad     # validation runs under no_grad: forward only
        holder = []
        total = _FusedLoss.apply(preds, gt, off, max(counts), anchors, strides, self.num_classes,
                                 self.lambda_cls, self.lambda_dfl, self.reg_max, need_grad, holder)
        stats = self.last_stats = holder[0]
        nc = self.num_classes

        def check(host):
            _raise_on_stall(host[6], "yb_loss_fwd_bwd")
            _raise_on_bad_class(host[7], nc)
        # one asynchronous D2H copy, waited for by the first read of a value (the reference does three .item())
        return total, LossDict(stats, ("total_loss", "box_loss", "cls_loss"), check)


# ----------------------------------------------------------------------------------------------
# public helpers of the reference module
# ----------------------------------------------------------------------------------------------
class _BboxIou(torch.autograd.Function):
    @staticmethod
    def forward(ctx, box1, box2):
        _cabi.require_cuda(box1, "box1")
        _cabi.require_cuda(box2, "box2")
        b1 = box1.detach().float().contiguous()
        b2 = box2.detach().float().contiguous()
        if b1.dim() != 2 or b1.shape[1] != 4 or b1.shape != b2.shape:
            raise ValueError(f"bbox_iou expects two (M, 4) tensors, got {tuple(box1.shape)} and {tuple(box2.shape)}")
        out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
        if b1.shape[0]:
            with torch.cuda.device(b1.device):
                rc = _cabi.lib().yb_bbox_iou(_cabi.ptr(b1), _cabi.ptr(b2), b1.shape[0], _cabi.ptr(out), None, None,
                                             _cabi.stream_ptr(b1.device))
            _cabi.check(rc, "yb_bbox_iou")
        ctx.save_for_backward(b1, b2)
        ctx.in_dtype = box1.dtype
        return out

    @staticmethod
    def backward(ctx, go):
        b1, b2 = ctx.saved_tensors
        g1 = torch.zeros_like(b1)
        if b1.shape[0]:
            go = go.detach().float().contiguous()
            scratch = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
            with torch.cuda.device(b1.device):
                rc = _cabi.lib().yb_bbox_iou(_cabi.ptr(b1), _cabi.ptr(b2), b1.shape[0], _cabi.ptr(scratch),
                                             _cabi.ptr(go), _cabi.ptr(g1), _cabi.stream_ptr(b1.device))
            _cabi.check(rc, "yb_bbox_iou")
        return g1.to(ctx.in_dtype), None


def bbox_iou(box1, box2):
    """Element-wise IoU of (M, 4) xywh pairs -> (M,), eps 1e-6; differentiable w.r.t. ``box1``.

    Reference ``bbox_iou`` (losses.py:9-40) including its ``b1_y2 = h + cy/2`` (line 20)."""
    return _BboxIou.apply(box1, box2)


class _Qfl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_scores, target_scores, beta):
        _cabi.require_cuda(pred_scores, "pred_scores")
        _cabi.require_cuda(target_scores, "target_scores")
        if not beta > 0:
            raise ValueError("quality_focal_loss: beta must be positive")
        if pred_scores.dim() != 2 or pred_scores.shape != target_scores.shape:
            raise ValueError("quality_focal_loss expects two (M, C) tensors of the same shape")
        x = pred_scores.detach().float().contiguous()
        t = target_scores.detach().float().contiguous()
        m, c = x.shape
        lib = _cabi.lib()
        ws = _workspace(lib.yb_qfl_workspace_bytes(x.numel()), x.device)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if pred_scores.requires_grad else None
        with torch.cuda.device(x.device):
            rc = lib.yb_quality_focal_loss(_cabi.ptr(x), _cabi.ptr(t), m, c, float(beta), _cabi.ptr(out),
                                           _cabi.ptr(grad), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(x.device))
        _cabi.check(rc, "yb_quality_focal_loss")
        ctx.grad = grad
        ctx.in_dtype = pred_scores.dtype
        return out[0]

    @staticmethod
    def backward(ctx, go):
        g = ctx.grad
        ctx.grad = None
        return (None if g is None else (g * go).to(ctx.in_dtype)), None, None


def quality_focal_loss(pred_scores, target_scores, beta=2.0):
    """Reference ``quality_focal_loss`` (losses.py:46-57): -sum(...)/M over dense (M, C) tensors."""
    return _Qfl.apply(pred_scores, target_scores, beta)


class _Dfl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_dist, target_val):
        _cabi.require_cuda(pred_dist, "pred_dist")
        _cabi.require_cuda(target_val, "target_val")
        if pred_dist.dim() != 2 or target_val.dim() != 1 or pred_dist.shape[0] != target_val.shape[0]:
            raise ValueError("distribution_focal_loss expects pred_dist (M, R) and target_val (M,)")
        x = pred_dist.detach().float().contiguous()
        t = target_val.detach().float().contiguous()
        m, r = x.shape
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if pred_dist.requires_grad else None
        with torch.cuda.device(x.device):
            rc = _cabi.lib().yb_distribution_focal_loss(_cabi.ptr(x), _cabi.ptr(t), m, r, _cabi.ptr(out),
                                                        _cabi.ptr(grad), _cabi.stream_ptr(x.device))
        _cabi.check(rc, "yb_distribution_focal_loss")
        ctx.grad = grad
        ctx.in_dtype = pred_dist.dtype
        return out[0]

    @staticmethod
    def backward(ctx, go):
        g = ctx.grad
        ctx.grad = None
        return (None if g is None else (g * go).to(ctx.in_dtype)), None


def distribution_focal_loss(pred_dist, target_val):
    """Reference ``distribution_focal_loss`` (losses.py:63-78): mean over M of the left/right CE."""
    return _Dfl.apply(pred_dist, target_val)
