"""ctypes binding of ``csrc/libyolo_boxpath.so`` (the C ABI declared in ``include/yolo_boxpath.h``).

There is no fallback of any kind: if the library has not been built, or a tensor is not on a CUDA
device, the call raises.  PyTorch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_double, c_float, c_int, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libyolo_boxpath.so")
ABI_VERSION = 6
YB_F32, YB_BF16 = 0, 1
YB_LOSS_NO_PRUNE, YB_LOSS_SPLIT_LAUNCH, YB_LOSS_FORCE_PROBE, YB_LOSS_WS_CLEAN = 1, 2, 4, 8
YB_LOSS_NO_PDL = 16

_lib = None
_lock = threading.Lock()

# ``_cabi.launch_count`` (module attribute, see __getattr__ below): kernels the library has launched in
# this process, counted inside the library at every launch site (bench.py reports the delta as gpu_launches)

_SIGNATURES = {
    "yb_abi_version": (c_int, []),
    "yb_last_error": (ctypes.c_char_p, []),
    "yb_struct_size": (c_size_t, [c_int]),
    "yb_launch_count": (ctypes.c_longlong, []),
    "yb_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "yb_loss_fwd_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, ctypes.c_uint, c_void_p, c_void_p, c_void_p]),
    "yb_tal_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "yb_tal_assign": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_tal_loss": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_peer_mailbox_bytes": (c_size_t, []),
    "yb_peer_mailbox_alloc": (c_int, [c_void_p]),
    "yb_peer_mailbox_free": (c_int, [c_void_p]),
    "yb_peer_mailbox_export": (c_int, [c_void_p, c_void_p]),
    "yb_peer_mailbox_open": (c_int, [c_void_p, c_void_p]),
    "yb_peer_mailbox_close": (c_int, [c_void_p]),
    "yb_scale_grad": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "yb_loss_fwd_bwd_host": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_dfl_decode": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_void_p]),
    "yb_make_anchors": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "yb_head_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yb_head_scatter": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "yb_dist2bbox": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "yb_val_decode_workspace_bytes": (c_size_t, [c_int, c_int]),
    "yb_val_decode": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "yb_nms_multilabel_workspace_bytes": (c_size_t, [c_int]),
    "yb_nms_multilabel": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_double, c_int, c_int, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_double, c_int, c_int, c_void_p, c_int, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_postprocess_workspace_bytes": (c_size_t, [c_int, c_int]),
    "yb_postprocess": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_double,
                               c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "yb_gather_gt": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "yb_xywh2xyxy": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "yb_bbox_iou": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "yb_box_iou": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "yb_box_iou_batch": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "yb_detection_counters_bytes": (c_size_t, [c_int]),
    "yb_detection_match": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_float, c_int, c_void_p, c_void_p]),
    "yb_qfl_workspace_bytes": (c_size_t, [c_size_t]),
    "yb_quality_focal_loss": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "yb_distribution_focal_loss": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)


class TalParams(ctypes.Structure):
    """``yb_tal_params`` of include/yolo_boxpath.h."""
    _fields_ = [("topk", c_int), ("alpha", c_float), ("beta", c_float), ("lambda_box", c_float), ("lambda_cls", c_float),
                ("lambda_dfl", c_float), ("vfl", c_int), ("vfl_alpha", c_float), ("vfl_gamma", c_float), ("flags", ctypes.c_uint)]


YB_TAL_WS_CLEAN = 1


TAL_MAX_LEVELS = 8
PEER_MAX_WORLD = 16
PEER_HANDLE_BYTES = 64


class PeerExchangeStruct(ctypes.Structure):
    """``yb_peer_exchange`` of include/yolo_boxpath.h."""
    _fields_ = [("world", c_int), ("rank", c_int), ("seq", ctypes.c_uint), ("mailbox", c_void_p * PEER_MAX_WORLD)]



class TalGrid(ctypes.Structure):
    """``yb_tal_grid`` of include/yolo_boxpath.h."""
    _fields_ = [("n_levels", c_int), ("start", c_int * TAL_MAX_LEVELS), ("w", c_int * TAL_MAX_LEVELS), ("h", c_int * TAL_MAX_LEVELS),
                ("stride", c_float * TAL_MAX_LEVELS), ("x0", c_float * TAL_MAX_LEVELS), ("y0", c_float * TAL_MAX_LEVELS)]


class ExtensionMissing(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The loaded shared library; raises ExtensionMissing when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ExtensionMissing(
                        f"{LIB_PATH} is missing: build the CUDA extension with "
                        "`python custom-yolo-implmentation_b200/build.py` (there is no CPU fallback)")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)          # AttributeError if the ABI lost a symbol
                    fn.restype, fn.argtypes = res, args
                if handle.yb_abi_version() != ABI_VERSION:
                    raise ExtensionMissing(f"{LIB_PATH}: ABI version {handle.yb_abi_version()} != {ABI_VERSION}; rebuild")
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().yb_last_error().decode(errors='replace')}")


def ptr(t) -> c_void_p:
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def stream_ptr(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}; this package runs on CUDA devices only (no CPU fallback)")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return YB_F32
    if dt == torch.bfloat16:
        return YB_BF16
    raise TypeError(f"unsupported dtype {dt}; the kernels take float32 or bfloat16 head outputs")


def __getattr__(name):
    if name == "launch_count":
        return int(lib().yb_launch_count())
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
