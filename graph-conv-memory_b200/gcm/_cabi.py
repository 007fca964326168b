"""ctypes binding of libgcm_b200.so (C ABI declared in include/gcm_b200.h).

PyTorch is only the owner of device memory and streams here: every call passes raw
`data_ptr()`s, sizes and the current CUDA stream handle.  There is no CPU fallback: if the
library is missing, `lib()` raises and every fused entry point fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

GCM_ABI_VERSION = 1
GCM_MAX_HOPS = 16
GCM_MAX_SELECTORS = 4
GCM_MAX_N = 1024
GCM_MAX_FEAT = 256

FLAG_NONFINITE = 1
FLAG_UNCLEAN = 2
FLAG_BADCOUNT = 4
FLAG_OVERFLOW = 8
FLAG_NONCAUSAL = 16
FLAG_NOTDENSE = 32

SEL_NONE, SEL_TEMPORAL, SEL_DENSE, SEL_EUCLIDEAN, SEL_COSINE, SEL_SPATIAL = range(6)
DIR = {"forward": 0, "backward": 1, "both": 2}
ACT = {"none": 0, "tanh": 1, "relu": 2}
ACT_EXP2X = 3          # gcm_linear2 epilogue only: exp(2 clamp(z)), the tanh form of the ones path cache
TK_AUTO, TK_HC, TK_TC, TK_WIN, TK_ROWS = range(5)
EB_AUTO, EB_PAIRS, EB_HASH = range(3)
GC_AUTO, GC_CUDA_CORES, GC_TC = range(3)
STEP_PURE_TEMPORAL = 1
STEP_UNIFORM_COUNT = 2
STEP_HCACHE_VALID = 4
STEP_WEIGHTS_STABLE = 8
STEP_COUNT_SHIFT = 8

_LIB_NAME = "libgcm_b200.so"
_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib")


class DenseStateC(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p), ("masks", C.c_void_p), ("count", C.c_void_p),
        ("B", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("F", C.c_int32), ("W", C.c_int32),
    ]


class SelectorC(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("direction", C.c_int32), ("n_hops", C.c_int32),
        ("hops", C.c_int32 * GCM_MAX_HOPS),
        ("max_distance", C.c_float),
        ("a_start", C.c_int32), ("a_step", C.c_int32), ("b_start", C.c_int32), ("b_step", C.c_int32),
        ("slice_len", C.c_int32),
        ("dist_param", C.c_void_p), ("dist", C.c_void_p),
    ]


class GnnC(C.Structure):
    _fields_ = [
        ("w1t", C.c_void_p), ("b1", C.c_void_p), ("w2t", C.c_void_p), ("b2", C.c_void_p),
        ("w_rel1", C.c_void_p), ("w_root1", C.c_void_p), ("w_rel2", C.c_void_p), ("w_root2", C.c_void_p),
        ("F", C.c_int32), ("H1", C.c_int32), ("H2", C.c_int32), ("act1", C.c_int32), ("act2", C.c_int32),
    ]


class RolloutC(C.Structure):
    """gcm_rollout (include/gcm_b200.h): what the host keeps between the steps of a temporal rollout."""
    _fields_ = [
        ("st", DenseStateC), ("gnn", GnnC), ("sels", SelectorC * GCM_MAX_SELECTORS),
        ("n_sels", C.c_int32), ("max_hop", C.c_int32),
        ("hcache", C.c_void_p), ("hc_ring", C.c_int32),
        ("uniform_count", C.c_int32), ("hc_fresh", C.c_int32), ("weights_stable", C.c_int32),
        ("status", C.c_void_p), ("scratch_obs", C.c_void_p), ("scratch_belief", C.c_void_p),
        ("launches", C.c_longlong),
        ("xrec", C.c_void_p), ("xrec_row0", C.c_longlong), ("xrec_from", C.c_int32),
    ]


class GnnGradsC(C.Structure):
    _fields_ = [
        ("d_w_rel1", C.c_void_p), ("d_w_root1", C.c_void_p), ("d_b1", C.c_void_p),
        ("d_w_rel2", C.c_void_p), ("d_w_root2", C.c_void_p), ("d_b2", C.c_void_p),
    ]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_SIGNATURES = {
    "gcm_version": (C.c_int, []),
    "gcm_last_error": (C.c_char_p, []),
    "gcm_last_kernel": (C.c_char_p, []),
    "gcm_launch_count": (C.c_longlong, []),
    "gcm_dense_step_fwd": (_I, [C.POINTER(DenseStateC), _P, C.POINTER(SelectorC), _I, C.POINTER(GnnC),
                                _P, _P, _I, _P]),
    "gcm_dense_step_fwd_cached": (_I, [C.POINTER(DenseStateC), _P, C.POINTER(SelectorC), _I, C.POINTER(GnnC),
                                       _P, _P, _I, _P, _I, C.POINTER(C.c_int), _P]),
    "gcm_dense_step_fwd_ex": (_I, [C.POINTER(DenseStateC), _P, _L, C.POINTER(SelectorC), _I, C.POINTER(GnnC),
                                   _P, _L, _P, _I, _I, _P, _I, C.POINTER(C.c_int), _P]),
    "gcm_dense_rollout_fwd": (_I, [C.POINTER(RolloutC), _P, _L, _L, _P, _L, _L, _I, _P]),
    "gcm_dense_rollout_step": (_I, [C.POINTER(RolloutC), _P, _P, _P]),
    "gcm_state_log_write_seq": (_I, [C.POINTER(DenseStateC), _P, _L, _L, _I, _P]),
    "gcm_temporal_gather": (_I, [C.POINTER(DenseStateC), _P, _I, _L, _I, _P, _I, _P]),
    "gcm_temporal_shift_sum": (_I, [_P, _L, _I, _L, _P, _I, _I, _P, _L, _I, _I, _I, _I, _P, _I, _P]),
    "gcm_temporal_shift_sum_strided": (_I, [_P, _L, _L, _L, _L, _I, _L, _P, _I, _I, _P, _L, _I, _I, _I, _I, _P, _I, _P]),
    "gcm_temporal_window_bwd_workspace": (_L, []),
    "gcm_temporal_window_bwd_set_trace": (_I, [_P, _I]),
    "gcm_temporal_window_bwd": (_I, [_P, _P, _L, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "gcm_set_temporal_kernel": (_I, [_I]),
    "gcm_dense_step_bwd": (_I, [C.POINTER(DenseStateC), _I, C.POINTER(GnnC), _P, _P, _P,
                                C.POINTER(GnnGradsC), _P]),
    "gcm_state_materialize": (_I, [C.POINTER(DenseStateC), _P, _P, _P, _P]),
    "gcm_state_materialize_grad": (_I, [C.POINTER(DenseStateC), _P, _P, _P]),
    "gcm_state_log_write": (_I, [C.POINTER(DenseStateC), _P, _I, _P]),
    "gcm_state_ingest": (_I, [C.POINTER(DenseStateC), _P, _P, _P, _P, _P]),
    "gcm_euclid_batchmean": (_I, [C.POINTER(DenseStateC), _P, _I, _P, _P, _P]),
    "gcm_euclid_tc_scratch": (C.c_longlong, [_I, _I]),
    "gcm_euclid_batchmean_tc": (_I, [C.POINTER(DenseStateC), _P, _I, _P, _P, _P, _P]),
    "gcm_select_dense": (_I, [_P, _P, _P, _I, _I, _I, C.POINTER(SelectorC), _P]),
    "gcm_sparse_write_flatten": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "gcm_sparse_write_flatten_oop": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "gcm_sparse_write_flatten_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "gcm_sparse_build_edges": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, C.c_float, _P,
                                    _P, _P, _L, _P, _P, _P, _I, _P]),
    "gcm_sparse_expand_edges": (_I, [_P, _P, _P, _I, _I, _P, _I, _P, _P, _L, _P, _P, _P]),
    "gcm_dense_ones_update": (_I, [C.POINTER(DenseStateC), _P, _P, _P, _P]),
    "gcm_dense_ones_xsum": (_I, [C.POINTER(DenseStateC), _P, _P]),
    "gcm_linear2": (_I, [_P, _I, C.c_longlong, _P, _P, _I, C.c_longlong, _P, _P, _I, C.c_longlong, _I, _P,
                         C.c_longlong, _P, _I, _P]),
    "gcm_outer_reduce": (_I, [_P, C.c_longlong, _I, _P, C.c_longlong, _I, C.c_longlong, _P, _P, _P]),
    "gcm_dense_ones_fwd": (_I, [C.POINTER(DenseStateC), _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "gcm_dense_ones_window_bwd": (_I, [C.POINTER(DenseStateC), _I, _I, _I, _P, _I, _I, _P, _P, _P, C.c_longlong, _P, _P]),
    "gcm_dense_ones_node_bwd": (_I, [C.POINTER(DenseStateC), _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, C.c_longlong, _P,
                                     _P]),
    "gcm_dense_ones_seq_update": (_I, [C.POINTER(DenseStateC), _P, C.c_longlong, C.c_longlong, _I, _P, _P, C.c_longlong, _P]),
    "gcm_dense_ones_seq_smem": (C.c_longlong, [_I, _I, _I, _I]),
    "gcm_dense_ones_window_fwd": (_I, [C.POINTER(DenseStateC), _I, _I, _I, _P, _I, _I, _P, _P, C.c_longlong, _P, _P, _P,
                                       C.c_longlong, _P]),
    "gcm_act_backward": (_I, [_P, _P, _I, C.c_longlong, _P, _P]),
    "gcm_act_backward_strided": (_I, [_P, _L, _L, _P, _I, _I, _I, _I, _P, _P]),
    "gcm_dense_ones_dc": (_I, [_P, _P, _P, _P, _I, C.c_longlong, _P, _P, _P, _P]),
    "gcm_to_bf16": (_I, [_P, _P, C.c_longlong, _P]),
    "gcm_linear_tc": (_I, [_P, _I, C.c_longlong, _P, _P, _I, C.c_longlong, _I, _P, C.c_longlong, _I, _P]),
    "gcm_linear_tc32": (_I, [_P, _I, C.c_longlong, _P, _P, _I, C.c_longlong, _P, _P, _I, C.c_longlong, _I, _P,
                             C.c_longlong, _P, _P]),
    "gcm_outer_reduce_tc_workspace": (C.c_longlong, [C.c_longlong]),
    "gcm_outer_reduce_tc": (_I, [_P, C.c_longlong, _I, _P, C.c_longlong, _I, C.c_longlong, _P, _P, _P, _P]),
    "gcm_outer_reduce_tc32": (_I, [_P, C.c_longlong, _I, _P, C.c_longlong, _I, C.c_longlong, _P, _P, _P, _P]),
    "gcm_outer_reduce_tc32_pair": (_I, [_P, C.c_longlong, _I, _P, C.c_longlong, _I, _P, C.c_longlong, _I, C.c_longlong, _P,
                                        _P, _P, _P, _P]),
    "gcm_dense_fill_masks": (_I, [C.POINTER(DenseStateC), _P]),
    "gcm_dense_step_fwd_zc": (_I, [C.POINTER(DenseStateC), _P, C.POINTER(SelectorC), C.POINTER(GnnC), _P, _P, _P, _P]),
    "gcm_set_edge_builder": (_I, [_I]),
    "gcm_set_graphconv_kernel": (_I, [_I]),
    "gcm_sparse_graphconv_hint_rows": (_I, [C.c_longlong]),
    "gcm_sparse_graphconv_hint_blocks": (_I, [_P, _I, _I]),
    "gcm_sparse_graphconv_fwd": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P, _P, _I, _P, _P, _P]),
    "gcm_sparse_csr_transpose": (_I, [_P, _P, _P, _P, _I, _L, _P, _P, _P]),
    "gcm_set_csr_transpose_cap": (_I, [_I]),
    "gcm_sparse_graphconv_bwd": (_I, [_P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P,
                                      _P, _P, _P, _P, _P, _P, _P, _P]),
    "gcm_pack_edges": (_I, [_P, _P, _L, _I, _I, _L, C.c_float, _P, _P, _P, _P]),
    "gcm_count_valid_edges": (_I, [_P, _I, _I, _P, _P]),
    "gcm_unpack_edges": (_I, [_P, _P, _I, _I, _P, _L, _P, _P, _P]),
    "gcm_any_nonfinite": (_I, [_P, _L, _P, _P]),
    "gcm_tc_selftest": (_I, [_P, _P, _P, _I, _I, _I, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


class GcmLibraryError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("GCM_B200_LIB", os.path.join(_LIB_DIR, _LIB_NAME))


def lib() -> C.CDLL:
    """Load the shared library once.  Raises GcmLibraryError if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise GcmLibraryError(
            f"{path} not found: the GCM hot path has no CPU or eager fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)."
        )
    handle = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise GcmLibraryError(f"{path} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    v = handle.gcm_version()
    if v != GCM_ABI_VERSION:
        raise GcmLibraryError(f"{path}: ABI version {v}, expected {GCM_ABI_VERSION}; rebuild")
    _lib = handle
    return handle


ERR_UNSUPPORTED = -3

_TRACE = bool(os.environ.get("GCM_B200_TRACE"))
_trace_t = [0.0]


def check(rc: int, what: str) -> None:
    if _TRACE:
        # debugging aid: synchronise after every library call and print what ran and how long it took
        import sys
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        print(f"[gcm trace] {what}: {lib().gcm_last_kernel().decode()} +{(now - _trace_t[0]) * 1e3:.2f} ms",
              file=sys.stderr, flush=True)
        _trace_t[0] = now
    if rc != 0:
        msg = lib().gcm_last_error().decode("utf-8", "replace")
        raise GcmLibraryError(f"{what} failed (status {rc}): {msg}")


def stream_ptr(device) -> int:
    """Raw cudaStream_t of torch's current stream on `device` (the kernels are enqueued there)."""
    idx = device.index if isinstance(device, torch.device) else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    try:
        return torch._C._cuda_getCurrentRawStream(idx)
    except AttributeError:      # older / newer torch without the private accessor
        return torch.cuda.current_stream(idx).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise GcmLibraryError(
            f"{what}: tensor is on {t.device}; the fused GCM path runs on CUDA (sm_100a) only and "
            "has no CPU fallback"
        )
