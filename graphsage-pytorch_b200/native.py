"""ctypes binding of libgsage_b200.so (the C ABI of include/gsage_b200.h).

This is the only place the Python host touches native code.  There is no fallback: if the
library cannot be loaded, or a call returns non-zero, a RuntimeError is raised.  Tensors are
passed as raw device pointers; every launch goes to torch's current CUDA stream so the calls
compose with torch ops and can be captured into CUDA graphs.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from typing import Optional

import torch

from . import build as _build

AGG_MEAN, AGG_MAX = 0, 1
SELF_KEEP, SELF_DROP, SELF_ONCE = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
MAX_FANOUT = 32
UNIQUE_MARKED, UNIQUE_LEAVE_MARKS = 1, 2
ABI_VERSION = 16

_P, _I, _L, _F, _U64, _SZ = c_void_p, c_int32, c_int64, c_float, c_uint64, c_size_t

# name -> (restype, argtypes); mirrors include/gsage_b200.h one to one
_SIGNATURES = {
    "gs_version": (_I, []),
    "gs_error_string": (ctypes.c_char_p, [_I]),
    "gs_launch_count": (_L, []),
    "gs_launch_count_reset": (None, []),
    "gs_set_pdl": (None, [_I]),
    "gs_sample_neighbors": (_I, [_P, _P, _L, _P, _P, _I, _I, _I, _I, _U64, _U64, _P, _P, _P, _P]),
    "gs_sample_neighbors_ex": (_I, [_P, _P, _L, _P, _P, _I, _I, _I, _I, _U64, _U64, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P]),
    "gs_fetch_batch": (_I, [_P, _I, _P, _P]),
    "gs_unique_remap_bitmap_ex": (_I, [_P, _P, _I, _P, _I, _L, _P, _P, _P, _P, _P, _SZ, _I, _P]),
    "gs_unique_workspace_bytes": (_SZ, [_I, _I]),
    "gs_unique_remap": (_I, [_P, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "gs_unique_bitmap_workspace_bytes": (_SZ, [_L]),
    "gs_unique_remap_bitmap": (_I, [_P, _P, _I, _P, _I, _L, _P, _P, _P, _P, _P, _SZ, _P]),
    "gs_agg_fwd": (_I, [_P, _L, _I, _P, _I, _P, _P, _I, _I, _P, _L, _P, _L, _P]),
    "gs_debug_stamp": (_I, [_P, _P]),
    "gs_set_agg_ctas": (None, [_I]),
    "gs_set_background": (None, [_I]),
    "gs_set_early_reads": (None, [_I]),
    "gs_agg_bwd": (_I, [_P, _L, _P, _L, _I, _P, _I, _P, _P, _P, _L, _P, _I, _I, _P, _L, _P, _L, _P]),
    "gs_sage_gemm_fwd": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _I, _I, _P, _I, _P, _L, _I, _I, _P]),
    "gs_sage_gemm_fwd_ex": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _I, _I, _P, _I, _P, _L, _I, _I, _P, _L, _P, _P, _I, _P]),
    "gs_agg_fwd_x": (_I, [_P, _L, _I, _P, _I, _P, _P, _P, _I, _I, _P, _L, _I, _P, _P]),
    "gs_split_lo": (_I, [_P, _P, _L, _P]),
    "gs_sage_top_workspace_bytes": (_SZ, []),
    "gs_sage_top_sup": (_I, [_P, _L, _P, _I, _P, _P, _P, _I, _P, _L, _I, _I, _I, _P, _P, _I, _P, _P, _P, _L, _P, _L, _P, _L,
                             _P, _L, _P, _P, _P, _P, _P, _L, _P, _SZ, _I, _P, _P, _I, _P]),
    "gs_sage_gemm_bwd_w": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _P, _L, _I, _I, _I, _P, _I, _P, _L, _I, _P]),
    "gs_sage_gemm_bwd_w_group": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "gs_sage_gemm_bwd_w_pair": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _I, _P]),
    "gs_sage_gemm_bwd_x": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _P, _I, _P, _L, _P, _L, _I, _P]),
    "gs_relu_bwd_inplace": (_I, [_P, _L, _P, _L, _I, _P, _I, _P]),
    "gs_cls_fwd": (_I, [_P, _L, _I, _I, _P, _P, _I, _P, _I, _P]),
    "gs_cls_bwd": (_I, [_P, _P, _P, _L, _I, _I, _P, _I, _P, _L, _P, _P, _P, _I, _P]),
    "gs_nll_fwd_bwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "gs_cls_nll_fwd_bwd": (_I, [_P, _L, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _I, _P, _I, _P]),
    "gs_clip_sgd": (_I, [_P, _P, _P, _I, _L, _F, _F, _F, _I, _P, _P]),
    "gs_agg_fwd_bf16_sharded": (_I, [_P, _I, _L, _L, _I, _P, _I, _P, _P, _P, _I, _P, _L, _P, _L, _I, _P]),
    "gs_dp_state_bytes": (_SZ, []),
    "gs_dp_region_bytes": (_SZ, [_L, _I]),
    "gs_dp_region_recv_offset": (_SZ, []),
    "gs_dp_allreduce_clip_sgd": (_I, [_P, _L, _P, _I, _I, _P, _P, _P, _P, _I, _F, _F, _P, _U64, _P, _P, _P, _P, _P, _P]),
    "gs_dp_status": (_I, [_P, _P, _P, _P, _P]),
    "gs_peer_alloc": (_I, [_SZ, _P]),
    "gs_peer_free": (_I, [_P]),
    "gs_peer_export": (_I, [_P, _P]),
    "gs_peer_open": (_I, [_P, _P]),
    "gs_peer_close": (_I, [_P]),
    "gs_random_walk_pos": (_I, [_P, _P, _L, _P, _I, _I, _I, _P, _U64, _U64, _P, _P, _P]),
    "gs_negative_workspace_bytes": (_SZ, [_L, _I]),
    "gs_negative_sample": (_I, [_P, _P, _L, _P, _I, _I, _I, _P, _I, _U64, _U64, _P, _P, _P, _P, _SZ, _P]),
    "gs_negative_sample_ex": (_I, [_P, _P, _L, _P, _I, _I, _I, _P, _I, _P, _U64, _U64, _P, _P, _P, _P, _SZ, _P]),
    "gs_pair_loss_fwd": (_I, [_P, _L, _I, _P, _I, _P, _P, _P, _P, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "gs_pair_loss_bwd": (_I, [_P, _L, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
}

_lib: Optional[ctypes.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if the in-tree .so is absent or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    override = os.environ.get("GSAGE_LIB")             # diagnostics: a trace / probe build of the same sources
    if override:
        path, build_if_missing = override, False
    if build_if_missing and not _build.is_current():
        try:
            _build.build()
        except Exception as exc:                       # stale-but-present library is still usable
            if not os.path.exists(path):
                raise RuntimeError(f"libgsage_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(path):
        raise RuntimeError("libgsage_b200.so is missing; run `python graphsage-pytorch_b200/build.py` "
                           "(there is no CPU or PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)                        # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gs_version() != ABI_VERSION:
        raise RuntimeError(f"libgsage_b200.so ABI {lib.gs_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


_timeline = None      # (device int64 buffer, [(label, stream id)]) while timeline markers are being recorded


def timeline_begin(device, capacity: int = 512) -> None:
    """Diagnostics: from now on every checked launch is followed by a gs_debug_stamp marker on its stream.
    Capture/replay a step, then timeline_read()."""
    global _timeline
    _timeline = (torch.zeros((capacity,), dtype=torch.int64, device=device), [])


def timeline_mark(label: str) -> None:
    if _timeline is None:
        return
    buf, labels = _timeline
    if len(labels) >= buf.numel():
        return
    st = torch.cuda.current_stream()
    load().gs_debug_stamp(buf.data_ptr() + 8 * len(labels), st.cuda_stream)
    labels.append((label, st.cuda_stream))


def timeline_read():
    """[(label, stream id, ns)] in launch order; stops recording."""
    global _timeline
    if _timeline is None:
        return []
    buf, labels = _timeline
    _timeline = None
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    return [(lab, sid, t[i]) for i, (lab, sid) in enumerate(labels)]


def set_agg_ctas(ctas_per_sm: int) -> None:
    """Occupancy cap of the K3 forward grid for subsequent launches (0 = full); see gs_set_agg_ctas."""
    load().gs_set_agg_ctas(int(ctas_per_sm))


def set_background(on: bool) -> None:
    """Mark subsequent launches as background work of a two-branch step (see gs_set_background)."""
    load().gs_set_background(int(bool(on)))


def set_early_reads(on: bool) -> None:
    """Let subsequent launches read their index inputs before the PDL wait (see gs_set_early_reads for when that is
    correct: the inputs must come from another stream joined by an event, not from the launch chain itself)."""
    load().gs_set_early_reads(int(bool(on)))


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().gs_error_string(code)
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")
    if _timeline is not None:
        timeline_mark(what)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: the gsage_b200 hot path has no CPU fallback")


def launch_count() -> int:
    return int(load().gs_launch_count())


def launch_count_reset() -> None:
    load().gs_launch_count_reset()


def set_pdl(mode: int) -> None:
    """1 = programmatic dependent launch on, 0 = off, -1 = follow GS_PDL (see include/gsage_b200.h)."""
    load().gs_set_pdl(int(mode))
