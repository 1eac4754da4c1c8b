#!/usr/bin/env python
"""Benchmark of the box-geometry hot path (BASELINE.json metric: images/s through
decode+assign+loss(+bwd) and through NMS; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the fused decode + assign + DFL/QFL loss + backward over one batch of
synthetic head outputs.  Workload at every N: cfg2 of BASELINE.json per GPU (batch 128, 640x640,
8400 anchors, 80 classes, <=100 GT/img, fp32) — i.e. cfg3 (global batch 1024) at N=8: weak scaling.

  value       whole-job images/s with inputs resident in HBM, timed with CUDA events on the launching
              stream, max over ranks (inputs are 619 MB/step: larger than the 126 MB L2).
  e2e         the same metric through the public drop-in API (YoloDFLQFLoss.forward + backward) with
              HOST inputs: every step copies preds / GT from pinned host memory and reads the loss
              scalars back.
  roofline    dominant kernel (fused_main_kernel): algorithmic bytes / CUDA-event duration vs the
              measured copy bandwidth in MEASURED_PEAKS.json.
  parity      on-hardware correctness of THIS run: every rank's loss scalars and matched anchors against the
              reference's own outputs for that rank's batch (tests/golden/loss_cfg3_ranks.npz).
  cpu_baseline / --impl reference
              the UNMODIFIED reference (baseline/_ref, vendored by baseline/vendor_ref.py) on the box's host
              cores on a bounded sample; falls back to the oracle port only when baseline/_ref is missing.
  cuda_eager_baseline (N=1)
              the same unmodified reference code on device="cuda" (PyTorch eager on this B200): what one
              gets without this repo.
  nms         cfg4 (batch 64, 8400 candidates, conf 0.001, IoU 0.7, max_det 300) through yb_nms, with the
              reference's non_max_suppression (torchvision.ops.nms per image) on CUDA and on the CPU beside it.
"""
import argparse
import json
import os
import statistics
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(batch_per_gpu=128, imgsz=640, nc=80, gmax=100, reg_max=16)
CPU_SAMPLE_IMAGES = 16
SEED0 = 1236                     # rank r draws its batch with SEED0 + r (tests/golden/loss_cfg3_ranks.npz pins ranks 0-7)
MIN_TIMED_STEPS = 100            # the timed window never shrinks below this (one NCCL call / barrier skew must not dominate)
METRIC = "images/sec through decode+assign+loss(+bwd)"
WORKLOAD = ("cfg2 per GPU: batch 128, 640x640 (8400 anchors, reg_max 16), 80 classes, <=100 GT/img, "
            "fused decode+assign+loss+backward (cfg3 = global batch 1024 at 8 GPUs)")


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class gpu_local_cpus:
    """Context manager: run on the CPUs NVML names as local to GPU `index` (first-touch places pinned host memory on
    that NUMA node; a staging buffer on the far socket halves the host-to-device rate).  Restores the affinity."""

    def __init__(self, index):
        self.index, self.old, self.note = index, None, "unchanged"

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.old = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.note = f"{len(os.sched_getaffinity(0))} of {len(self.old)} cpus"
        except Exception as e:                              # not permitted / not supported: leave the placement alone
            self.note = f"unavailable ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except OSError:
                pass
        return False


class ClockSampler:
    """SM clock / throttle reasons of one GPU through NVML, in-process.

    NVML queries take the driver lock for milliseconds and stall kernel launches, so nothing is
    sampled while the timed loop is ENQUEUEING; the samples are taken right after the last launch of
    the timed region, while the GPU is still draining the queued steps (i.e. under load, inside the
    timed region, before the closing synchronize)."""

    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows, self.h, self.mx, self.sample_ms = [], None, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                         pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
        except Exception:
            self.h = None

    def sample(self, n=1, gap=0.002):
        if self.h is None:
            return
        for i in range(n):
            t0 = time.perf_counter()
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), [nm for nm, b in zip(self.NAMES, self.bits) if r & b]))
            except Exception:
                pass
            self.sample_ms.append((time.perf_counter() - t0) * 1e3)
            if i + 1 < n:
                time.sleep(gap)

    def summary(self):
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.mx) if self.mx else None,
                "reasons": sorted({n for r in self.rows for n in r[1]}), "samples": len(sm),
                "nvml_ms_per_sample": round(statistics.mean(self.sample_ms), 3) if self.sample_ms else None}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# the unmodified reference (baseline/_ref), for the CPU / CUDA-eager legs.  Never on our arm's timed path.
# ---------------------------------------------------------------------------------------------
_REF = None


def reference_modules():
    """(losses, model_utils) of the vendored reference, or None when baseline/_ref is not there.  The only
    intervention is the frozen NMS clock, applied per call by ref_nms()."""
    global _REF
    if _REF is None:
        ref_root = os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isfile(os.path.join(ref_root, "src", "model", "losses.py")):
            _REF = False
        else:
            sys.path.insert(0, ref_root)
            try:
                import src.model.losses as ref_losses
                import src.utils.model_utils as ref_utils
                _REF = (ref_losses, ref_utils)
            except Exception as e:                       # e.g. a missing third-party import on the box
                sys.stderr.write(f"bench.py: baseline/_ref is not importable ({type(e).__name__}: {e})\n")
                _REF = False
            finally:
                sys.path.remove(ref_root)
    return _REF or None


def ref_matched_idx(preds, gts, anchors, strides, reg_max=16):
    """The matched anchors the reference computes but does not return: its own lines (losses.py:142-188 decode,
    :214-215 cdist + argmin) re-executed on the given device.  Concatenated over the images."""
    p = preds.float().transpose(1, 2)
    anc, st = anchors.transpose(0, 1), strides.transpose(0, 1)
    b, a, _ = p.shape
    pd = p[:, :, : 4 * reg_max].view(b, a, 4, reg_max).softmax(3)
    ltrb = torch.sum(pd * torch.arange(reg_max, device=p.device, dtype=p.dtype), dim=3)
    x1 = (anc[None, :, 0] - ltrb[:, :, 0]) * st[None, :, 0]
    y1 = (anc[None, :, 1] - ltrb[:, :, 1]) * st[None, :, 0]
    x2 = (anc[None, :, 0] + ltrb[:, :, 2]) * st[None, :, 0]
    y2 = (anc[None, :, 1] + ltrb[:, :, 3]) * st[None, :, 0]
    ctr = torch.stack([(x1 + x2) / 2, (y1 + y2) / 2], dim=2)
    out = [torch.cdist(g[:, 0:2].to(p.dtype), ctr[i]).argmin(dim=1) for i, g in enumerate(gts) if g.numel()]
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.long, device=p.device)


def ref_nms(ref_utils, y, frozen=True, **kw):
    """The reference's non_max_suppression; `frozen` stops its wall-clock abort (model_utils.py:212, :275-277)."""
    real = ref_utils.time
    if frozen:
        ref_utils.time = types.SimpleNamespace(time=lambda: 0.0)
    try:
        return ref_utils.non_max_suppression(y, **kw)
    finally:
        ref_utils.time = real


def cpu_loss_images_per_s(steps, warmup):
    """Reference loss forward + backward on the host cores over CPU_SAMPLE_IMAGES images of the cfg2 workload."""
    from custom_yolo_implmentation_b200.utils import synthetic as syn

    torch.set_num_threads(os.cpu_count() or 1)
    n = CPU_SAMPLE_IMAGES
    preds, gts, anchors, strides = syn.make_loss_inputs(n, CFG["nc"], CFG["imgsz"], CFG["gmax"], SEED0)
    ref = reference_modules()
    if ref is not None:
        kind, what = "reference", "the unmodified reference's YoloDFLQFLoss.forward + loss.backward (baseline/_ref/src/model/losses.py)"
        crit = ref[0].YoloDFLQFLoss(num_classes=CFG["nc"])

        def run():
            x = preds.clone().requires_grad_(True)
            loss, _ = crit(x, gts, anchors, strides)
            loss.backward()
    else:
        from oracle import loss_oracle
        kind, what = "port", "oracle/loss_oracle.py fwd+bwd (baseline/_ref missing)"

        def run():
            loss_oracle.loss_forward_backward(preds, gts, anchors, strides, CFG["nc"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return n / med, med, torch.get_num_threads(), kind, what


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    ips, med, threads, kind, what = cpu_loss_images_per_s(steps, warmup)
    sample = (f"{CPU_SAMPLE_IMAGES} images of the cfg2 workload per step (640x640, 8400 anchors, nc=80, <=100 GT/img, fp32), "
              f"{what} on torch CPU, median of {steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "cpu_sample_images": CPU_SAMPLE_IMAGES},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from custom_yolo_implmentation_b200 import _cabi
    from custom_yolo_implmentation_b200.model.losses import YoloDFLQFLoss, fused_loss, fused_tal_loss, pack_gt, pack_gt_host
    from custom_yolo_implmentation_b200.training.distributed_setup import reduce_loss_stats
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    from custom_yolo_implmentation_b200.utils.model_utils import batched_nms_raw, non_max_suppression

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()
    warmup = max(args.warmup, 3)
    timed_steps = max(args.steps, MIN_TIMED_STEPS)

    n, nc, imgsz, gmax = CFG["batch_per_gpu"], CFG["nc"], CFG["imgsz"], CFG["gmax"]
    preds_h, gts_h, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, SEED0 + rank)
    a = preds_h.shape[2]
    preds = preds_h.to(dev)
    anchors_d, strides_d = anchors.to(dev), strides.to(dev)
    gts_d = [g.to(dev) for g in gts_h]
    gt, off, counts = pack_gt(gts_d, dev)
    gmax_real = max(counts)
    bytes_per_step = 2 * preds.numel() * preds.element_size()          # read once + gradient written once
    peak, peak_src = measured_peak_gbs()

    def step():
        out, grad, _ = fused_loss(preds, gt, off, gmax_real, anchors_d, strides_d, nc, 1.0, 1.5, want_grad=True)
        return out, grad

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(warmup):                          # results bound exactly as in the timed loop, so the caching
        out, grad = step()                           # allocator already owns both 619 MB gradient blocks
    if world > 1:
        reduce_loss_stats(out, n)                    # NCCL connects its channels lazily on the first collective
    barrier()
    launches0 = _cabi.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    barrier()
    ev0.record()
    for _ in range(timed_steps):
        out, grad = step()
    if world > 1:
        # the path has no per-step exchange (every image is independent); like the reference, which
        # all-reduces its logging scalars once per epoch (src/training/train_model.py:285-288), the
        # loss statistics are reduced once per run — one 8-float message, inside the timed region
        reduced = reduce_loss_stats(out, n)
    ev1.record()
    clocks.sample(3)                         # the GPU is still draining the queued steps: samples under load
    barrier()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = (_cabi.launch_count - launches0) * args.steps // timed_steps      # per K steps, as the contract counts
    ms_per_step = elapsed_ms / timed_steps
    value = world * n * timed_steps / (elapsed_ms * 1e-3)
    loss_val = float(out[0].item())

    # ---- on-hardware parity of this very run: each rank against the reference's outputs for ITS batch ----
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", "loss_cfg3_ranks.npz")
    if os.path.exists(gpath) and rank < 8:
        import numpy as np
        z = np.load(gpath)
        if tuple(int(v) for v in z["meta"][1:]) == (n, nc, imgsz, gmax, SEED0):
            _, _, tr = fused_loss(preds, gt, off, gmax_real, anchors_d, strides_d, nc, 1.0, 1.5, want_grad=False, want_trace=True)
            host = out.cpu()
            rel = max(abs(host[k].item() - float(z[key][rank])) / abs(float(z[key][rank]))
                      for k, key in enumerate(("total_loss", "box_loss", "cls_loss")))
            idx_ref = torch.from_numpy(z["idx"][rank, : int(sum(counts))].astype("int64"))
            same = tr["idx"].cpu().long() == idx_ref
            agree = int(same.sum())
            # What a flipped exact tie may move: the GT goes to the other (equidistant) anchor and its own DFL term changes
            # -- at most ~24 (four sides, cross-entropy over 16 bins of O(1) logits) against a mean of ~3.3 per side -- with the
            # weight 1 / (4 M N) of a GT in an image of M boxes: 24 / (4 * 3.3 * N * M) = 1.5e-2 / M at N = 128, relative.
            img_of_gt = torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts))
            m_of_gt = torch.tensor(counts, dtype=torch.float64)[img_of_gt]
            allow = 1e-5 + float((1.5e-2 * (128.0 / n) / m_of_gt[~same]).sum())
            res = torch.tensor([rel, float(agree), float(sum(counts)), allow], dtype=torch.float64, device=dev)
            if world > 1:
                allr = [torch.zeros_like(res) for _ in range(world)]
                dist.all_gather(allr, res)
            else:
                allr = [res]
            allr = [r.cpu().tolist() for r in allr]
            parity = {"against": "the unmodified reference's loss scalars and matched anchors for each rank's batch "
                                 "(tests/golden/loss_cfg3_ranks.npz, generated by tests/golden/make_golden.py)",
                      "ranks_checked": len(allr), "loss_rel_err_max": max(r[0] for r in allr),
                      "loss_rel_err_per_rank": [r[0] for r in allr], "loss_rel_err_allowed_per_rank": [r[3] for r in allr],
                      "matched_anchors_identical": f"{int(sum(r[1] for r in allr))}/{int(sum(r[2] for r in allr))}"}
            # A GT whose two nearest predicted centres lie at the SAME float distance is decided by the last ulp of
            # the decode (rank 0's batch has four such exact ties; the reference's own CPU and CUDA runs differ on
            # them, see cuda_eager_baseline.matched_anchors_cuda_vs_cpu).  Ranks with every anchor identical must
            # meet 1e-5 on the loss; a flipped tie moves one GT's terms and nothing else.
            n_bad = sum(r[2] - r[1] for r in allr)
            parity["note"] = ("exact" if n_bad == 0 else
                              f"{int(n_bad)} GT(s) on an exact float tie of the distance went to the other (equidistant) anchor")
            # (bound per rank: 1e-5, plus 1.5e-2 / M for every flipped GT of an image with M boxes -- computed above, r[3])
            bad_rank = [i for i, r in enumerate(allr) if r[0] > r[3] or r[2] - r[1] > 1e-3 * r[2]]
            if bad_rank:
                raise SystemExit(f"bench.py: PARITY FAILURE on the benchmark batch (ranks {bad_rank}): {parity} {allr}")
    if world > 1:
        full = reduced.cpu()
        if not (abs(full[4].item() - world * n) < 0.5):
            raise SystemExit(f"bench.py: the reduced statistics count {full[4].item()} images, expected {world * n}")

    # ---- per-kernel durations for the roofline (separate pass, the caller's events on the launching stream) ----
    stage_ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    stage = []
    for _ in range(max(5, min(args.steps, 20))):
        fused_loss(preds, gt, off, gmax_real, anchors_d, strides_d, nc, 1.0, 1.5, want_grad=True, stage_events=stage_ev)
        torch.cuda.synchronize(dev)
        stage.append([stage_ev[0].elapsed_time(stage_ev[1]), stage_ev[1].elapsed_time(stage_ev[2])])
    isolated_ms, _ = (statistics.mean(s[i] for s in stage) for i in range(2))
    # The step is ONE launch and nothing else (no memset node: the launch's last CTA wipes its own counters), so the
    # launch duration the roofline uses is the timed region's: K back-to-back launches between two CUDA events / K.
    # (`ms_isolated_launch`: one launch onto an idle GPU between two events of its own, start-up latency included.)
    main_ms = ms_per_step
    # fused_main_kernel is the whole step: box role (reads the 64 box channels, writes their gradient) + class role
    # (reads the nc class channels, writes their gradient) + match role and reducer (per-GT terms, loss scalars):
    # every byte of preds read once, every byte of grad written once
    main_bytes = bytes_per_step
    main_gbs = main_bytes / (main_ms * 1e-3) / 1e9
    traffic = None                                   # dram read+write per launch from the committed ncu --set full capture
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath) and (n, nc, a, preds.element_size()) == (128, 80, 8400, 4):
            traffic = json.load(open(tpath)).get("fused_main_kernel")
            break
    roofline = {"bound": "hbm", "kernel": "fused_main_kernel", "achieved": main_gbs, "peak": peak, "unit": "GB/s",
                "frac": main_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": main_bytes, "ms_per_launch": main_ms, "ms_isolated_launch": isolated_ms,
                "other_kernels": {},                   # the step is this one launch
                "whole_step": {"algorithmic_bytes": bytes_per_step, "GB/s": bytes_per_step / (ms_per_step * 1e-3) / 1e9,
                               "frac": bytes_per_step / (ms_per_step * 1e-3) / 1e9 / peak}}

    def time_steps(fn, k):
        for _ in range(3):
            keep = fn()
        barrier()
        ev0.record()
        for _ in range(k):
            keep = fn()
        ev1.record()
        barrier()
        del keep
        return max_over_ranks(ev0.elapsed_time(ev1) / k)

    extra_steps = max(20, min(args.steps, 100))

    # ---- the same step through the public module, inputs already on the device ----
    # (what a training loop calls: GT list -> pack, autograd bridge, one D2H copy of the loss scalars per step)
    crit = YoloDFLQFLoss(num_classes=nc)
    x_leaf = preds.clone().requires_grad_(True)

    def api_step(gt_arg):
        x_leaf.grad = None
        loss, parts = crit(x_leaf, gt_arg, anchors_d, strides_d)
        loss.backward()
        return parts["total_loss"]               # read every step, as the reference's loop does after backward (train_model.py:255)

    packed_d = pack_gt_host(gts_h, pin_memory=False).to(dev)
    api_list_ms = time_steps(lambda: api_step(gts_d), extra_steps)
    api_packed_ms = time_steps(lambda: api_step(packed_d), extra_steps)
    api = {"what": "YoloDFLQFLoss.forward + loss.backward, head output and GT resident on the device",
           "gt_list_of_tensors": {"ms_per_step": api_list_ms, "value": world * n / (api_list_ms * 1e-3), "unit": "images/s",
                                  "note": "the reference's signature: 128 small device tensors per step, gathered by one yb_gather_gt launch from a host-written table; loss dict read every step"},
           "packed_gt": {"ms_per_step": api_packed_ms, "value": world * n / (api_packed_ms * 1e-3), "unit": "images/s",
                         "note": "PackedGT wire format (data/collate.py::collate_fn_packed); loss dict read every step"},
           "kernels_only_ms_per_step": ms_per_step}
    del x_leaf

    # ---- end to end through the public API, host inputs ----
    # Every step copies its own inputs from pinned host memory (the head output and the packed GT wire format of
    # data/collate.py::collate_fn_packed) and reads its loss scalars back.  As a data loader would, the copy of
    # step i+1 is issued on a copy stream while step i computes; all K copies lie inside the timed region.
    with gpu_local_cpus(local) as numa_note:           # pinned staging buffers on the GPU's own NUMA node
        preds_pin = preds_h.pin_memory()
        packed_pin = pack_gt_host(gts_h, pin_memory=True)
    h2d = preds_pin.numel() * preds_pin.element_size() + packed_pin.gt.numel() * 4 + packed_pin.offsets.numel() * 4
    d2h = 8 * 4
    copy_stream = torch.cuda.Stream(device=dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            x = preds_pin.to(dev, non_blocking=True)
            pg = packed_pin.to(dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        return x, pg, done

    def e2e_run(k):
        parts, nxt = None, fetch()
        for i in range(k):
            x, pg, done = nxt
            torch.cuda.current_stream().wait_event(done)
            if i + 1 < k:
                nxt = fetch()
            x.record_stream(torch.cuda.current_stream())
            pg.gt.record_stream(torch.cuda.current_stream())
            pg.offsets.record_stream(torch.cuda.current_stream())
            x.requires_grad_(True)
            loss, parts = crit(x, pg, anchors_d, strides_d)            # parts: one D2H copy of the loss scalars
            loss.backward()
            parts["total_loss"]                                          # read every step (waits for the step's D2H copy)
        return parts

    e2e_steps = max(3, min(args.steps, 10))
    parts = e2e_run(2)
    barrier()
    ev0.record()
    parts = e2e_run(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)
    del preds_pin

    # ---- NMS (cfg4), reported alongside ----
    y = syn.make_nms_input(64, nc, imgsz, 2024 + rank).to(dev)
    nms_ms = time_steps(lambda: batched_nms_raw(y, 0.001, 0.7, 300, nc), max(5, min(args.steps, 20)))
    rows, cnt, _ = batched_nms_raw(y, 0.001, 0.7, 300, nc)
    nms_bytes = y.numel() * 4 + 64 * 300 * 6 * 4
    nms = {"metric": "images/sec through batched class-aware NMS", "value": world * 64 / (nms_ms * 1e-3), "unit": "images/s",
           "ms_per_step": nms_ms, "config": {"workload": "cfg4: batch 64/GPU, 8400 candidates, nc=80, conf 0.001, IoU 0.7, max_det 300"},
           "hbm_frac_of_scan_roofline": nms_bytes / (nms_ms * 1e-3) / 1e9 / peak, "kept_min": int(cnt.min().item())}
    ppath = os.path.join(ROOT, "profiles", "r2_nms_pipes.json")          # from the committed ncu --set full capture
    if os.path.exists(ppath):
        nms["pipes_ncu"] = json.load(open(ppath))

    # ---- extra lines (not the headline): the task-aligned variant, the dense bf16 config, adverse class logits ----
    # the task-aligned path's one exchange step: [sum of target scores, #fg] averaged over the ranks between assignment and
    # loss.  Headline: through peer-mapped mailboxes (csrc/peer.cu: NVLink stores from the assign call's last kernel, a poll
    # in the loss call's first); beside it the same step with a torch.distributed (NCCL) all-reduce.
    px = None
    if world > 1:
        from custom_yolo_implmentation_b200.training.distributed_setup import PeerExchange
        px = PeerExchange()
    tal_ms = time_steps(lambda: fused_tal_loss(preds, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5, exchange=px), extra_steps)
    tal_out, _, _ = fused_tal_loss(preds, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5, exchange=px)
    tal = {"metric": "images/sec through decode+TAL assign+CIoU/DFL/BCE loss+bwd (no reference counterpart; parity vs in-repo oracle)",
           "value": world * n / (tal_ms * 1e-3), "unit": "images/s", "ms_per_step": tal_ms,
           "hbm_frac_whole_step": bytes_per_step / (tal_ms * 1e-3) / 1e9 / peak,
           "exchange": ("peer mailboxes over NVLink: [sum target scores, #fg] stored into every rank's mailbox by yb_tal_assign's last "
                        "kernel, polled by yb_tal_loss's first") if world > 1 else "none (1 GPU)"}
    if world > 1:                        # the exchange on real hardware: every rank must have used the SAME normaliser
        def normalisers(o):
            used = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(used, o[4:5].clone())
            return [float(u.item()) for u in used]
        used = normalisers(tal_out)
        nccl_ms = time_steps(lambda: fused_tal_loss(preds, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5, exchange=None), extra_steps)
        nccl_out, _, _ = fused_tal_loss(preds, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5, exchange=None)
        used_nccl = normalisers(nccl_out)
        tal["normaliser_used_by_rank"] = used
        tal["with_nccl_all_reduce_instead"] = {"ms_per_step": nccl_ms, "value": world * n / (nccl_ms * 1e-3), "unit": "images/s",
                                               "normaliser_used_by_rank": used_nccl}
        if max(used) != min(used) or max(used_nccl) != min(used_nccl) or abs(used[0] - used_nccl[0]) > 1e-6 * used[0]:
            raise SystemExit(f"bench.py: PARITY FAILURE: ranks normalised the task-aligned loss differently: {used} / {used_nccl}")
        px.close()
    # class logits ~ N(0, 2): half of them positive, nothing like the head's bias initialisation the headline input
    # follows; the class role's packed fast path (background logits <= -0.88) does not apply to most groups
    g = torch.Generator().manual_seed(4321 + rank)
    preds_w = preds_h.clone()
    preds_w[:, 64:] = torch.randn(preds_w[:, 64:].shape, generator=g) * 2.0
    preds_w = preds_w.to(dev)
    worst_ms = time_steps(lambda: fused_loss(preds_w, gt, off, gmax_real, anchors_d, strides_d, nc, 1.0, 1.5), extra_steps)
    worst_tal_ms = time_steps(lambda: fused_tal_loss(preds_w, gt, off, anchors_d, strides_d, nc, 1.5, 1.0, 1.5), extra_steps)
    worst = {"input": "cfg2 with class logits ~ N(0, 2) instead of N(-4.6, 1)", "ms_per_step": worst_ms,
             "value": world * n / (worst_ms * 1e-3), "unit": "images/s",
             "hbm_frac_whole_step": bytes_per_step / (worst_ms * 1e-3) / 1e9 / peak,
             "tal_ms_per_step": worst_tal_ms, "tal_hbm_frac_whole_step": bytes_per_step / (worst_tal_ms * 1e-3) / 1e9 / peak}
    del preds_w
    p5, g5, a5, s5 = syn.make_loss_inputs(32, nc, 1280, 300, 1240 + rank, dtype=torch.bfloat16)
    p5 = p5.to(dev); a5 = a5.float().to(dev); s5 = s5.float().to(dev)
    gt5, off5, c5 = pack_gt([g.to(dev) for g in g5], dev)
    cfg5_ms = time_steps(lambda: fused_loss(p5, gt5, off5, max(c5), a5, s5, nc, 1.0, 1.5), extra_steps)
    cfg5_bytes = 2 * p5.numel() * p5.element_size()
    cfg5 = {"config": "cfg5: batch 32/GPU, 1280x1280 (33600 anchors), <=300 GT/img, bf16 head outputs", "value": world * 32 / (cfg5_ms * 1e-3),
            "unit": "images/s", "ms_per_step": cfg5_ms, "hbm_frac_whole_step": cfg5_bytes / (cfg5_ms * 1e-3) / 1e9 / peak}
    del p5, gt5

    # ---- N = 1 only: the reference's own code on this box (CUDA eager, CPU), never part of our timed path ----
    cpu = cuda_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, med, threads, kind, what = cpu_loss_images_per_s(3, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
               "sample": f"{CPU_SAMPLE_IMAGES} images of the same workload per step, {what} on torch CPU, "
                         f"median of 3 steps ({med * 1e3:.0f} ms/step)"}
        ref = reference_modules()
        if ref is not None:
            ref_losses, ref_utils = ref
            rcrit = ref_losses.YoloDFLQFLoss(num_classes=nc)

            def ref_cuda_step():
                x = preds.detach().clone().requires_grad_(True)
                loss, parts = rcrit(x, gts_d, anchors_d, strides_d)
                loss.backward()
                return parts, x.grad

            ref_cuda_step()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(2):
                rparts, rgrad = ref_cuda_step()
            torch.cuda.synchronize(dev)
            ref_ms = (time.perf_counter() - t0) / 2 * 1e3
            gerr = float((rgrad - grad).abs().max().item() / rgrad.abs().max().item())
            cuda_vs_cpu = None
            if parity is not None:
                ridx = ref_matched_idx(preds, gts_d, anchors_d, strides_d).cpu()
                cuda_vs_cpu = f"{int((ridx == idx_ref).sum())}/{idx_ref.numel()}"
            cuda_eager = {"what": "the unmodified reference's YoloDFLQFLoss.forward + loss.backward on device='cuda' "
                                  "(PyTorch eager on this B200), full cfg2 batch", "ms_per_step": ref_ms,
                          "value": n / (ref_ms * 1e-3), "unit": "images/s", "speedup_of_this_repo": ref_ms / ms_per_step,
                          "speedup_through_public_api": ref_ms / api_list_ms,
                          "loss_rel_diff_vs_this_repo": abs(rparts["total_loss"] - loss_val) / abs(rparts["total_loss"]),
                          "grad_max_abs_diff_over_max": gerr,
                          "matched_anchors_cuda_vs_cpu": cuda_vs_cpu,     # the reference against itself: CUDA run vs the CPU golden
                          "note": "duplicate matched anchors make the reference's CUDA index_put_ order-dependent (SURVEY Q4)"}
            del rgrad
            # NMS: the reference's wrapper calls torchvision.ops.nms once per image
            kw = dict(conf_thres=0.001, iou_thres=0.7, max_det=300, nc=nc)
            ref_rows = ref_nms(ref_utils, y, **kw)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(3):
                ref_rows = ref_nms(ref_utils, y, **kw)
            torch.cuda.synchronize(dev)
            tv_ms = (time.perf_counter() - t0) / 3 * 1e3
            ours = non_max_suppression(y, **kw)
            same = all(torch.equal(o, r) for o, r in zip(ours, ref_rows))
            t0 = time.perf_counter()
            live = ref_nms(ref_utils, y, frozen=False, **kw)
            torch.cuda.synchronize(dev)
            live_ms = (time.perf_counter() - t0) * 1e3
            y_cpu = y[:4].cpu()
            ref_nms(ref_utils, y_cpu[:1], **kw)
            t0 = time.perf_counter()
            ref_nms(ref_utils, y_cpu, **kw)
            cpu_nms_ms = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            live_cpu = ref_nms(ref_utils, y.cpu(), frozen=False, **kw)
            live_cpu_s = time.perf_counter() - t0
            nms["reference_torchvision_cuda"] = {
                "what": "the unmodified reference's non_max_suppression on device='cuda' (torchvision.ops.nms per image, clock frozen), same 64 images",
                "ms_per_step": tv_ms, "value": 64 / (tv_ms * 1e-3), "unit": "images/s", "rows_identical_to_this_repo": bool(same),
                "unfrozen_clock": {"ms": live_ms, "images_left_empty": sum(1 for r in live if r.shape[0] == 0)}}
            nms["vs_torchvision_cuda"] = tv_ms / nms_ms
            nms["reference_cpu"] = {"what": "the same function on the host cores, clock frozen, 4 of the 64 images",
                                    "ms_per_image": cpu_nms_ms / 4, "value": 4 / (cpu_nms_ms * 1e-3), "unit": "images/s",
                                    "cores": torch.get_num_threads(),
                                    "unfrozen_clock_64_images": {"seconds": live_cpu_s, "images_left_empty": sum(1 for r in live_cpu if r.shape[0] == 0),
                                                                 "note": "the reference stops after 0.5 + 0.05 * 64 = 3.7 s (model_utils.py:212, :275-277)"}}
            if not same:
                raise SystemExit("bench.py: PARITY FAILURE: NMS rows differ from the reference's on the benchmark input")

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "timed_steps": timed_steps, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * n, "gt_boxes_per_step_per_gpu": int(sum(counts)),
                           "l2_policy": "inputs+outputs are 1.24 GB per step, larger than the 126 MB L2",
                           "timed_window": f"max(K, {MIN_TIMED_STEPS}) consecutive steps",
                           "parallelism": (f"batch-sharded x{world}: no data-path collective, one 8-float all-reduce of the loss statistics per run"
                                           if world > 1 else "single GPU")},
                "roofline": roofline, "cpu_baseline": cpu, "cuda_eager_baseline": cuda_eager, "parity": parity,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                        "api": "YoloDFLQFLoss.forward + loss.backward on pinned host inputs (head output + packed GT); the next step's copy is prefetched on a copy stream",
                        "pinned_numa": numa_note.note},
                "api_device_resident": api,
                "gpu_launches": launches, "clocks": clocks.summary(), "loss": loss_val, "nms": nms, "tal": tal,
                "cfg5_bf16": cfg5, "adverse_logits": worst}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
