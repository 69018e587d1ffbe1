"""ctypes binding of libnwk.so -- the C ABI declared in include/nwk.h.

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception.  The library is built in-tree by ``_build.build()`` (nvcc,
sm_100a) and loaded from this package directory so the driver sees which native code
served a run.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint32,
                    c_uint64, c_void_p)

from . import _build

METRIC_SQEUCLIDEAN, METRIC_COSINE, METRIC_DOT = 0, 1, 2
METRIC_IDS = {"sqeuclidean": METRIC_SQEUCLIDEAN, "cosine": METRIC_COSINE, "dot": METRIC_DOT}
FLAG_FORCE_EXACT_SCAN = 0x1
FLAG_NO_FALLBACK = 0x2
FLAG_PROFILE = 0x4
FLAG_TF32 = 0x8
FLAG_ALPHA_BETA = 0x10
MAX_K = 1024
ABI_VERSION = 2
MERGE_MAX_K = 2048

# every symbol include/nwk.h declares (tests/test_abi.py checks the library exports them all)
EXPORTED_SYMBOLS = (
    "nwk_last_error", "nwk_abi_version", "nwk_launch_count", "nwk_knn_workspace_bytes", "nwk_knn_f32",
    "nwk_merge_workspace_bytes", "nwk_merge_topk", "nwk_topk_state_create", "nwk_topk_state_update",
    "nwk_topk_state_search", "nwk_topk_state_finalize", "nwk_topk_state_destroy", "nwk_knn_host_f32", "nwk_synth_fill",
    "nwk_debug_tc_tile", "nwk_validate_pairs", "nwk_maxsim_workspace_bytes", "nwk_maxsim_f32",
    "nwk_merge_topk_strided", "nwk_merge_topk_ptrs", "nwk_topk_state_reset", "nwk_host_slab_plan", "nwk_host_release",
    "nwk_rows_from_columns_f32", "nwk_xchg_create", "nwk_xchg_handle_bytes", "nwk_xchg_local_handle",
    "nwk_xchg_open_peers", "nwk_xchg_next_slot", "nwk_xchg_merge", "nwk_xchg_check", "nwk_xchg_destroy",
    "nwk_host_alloc", "nwk_host_free", "nwk_mg_create", "nwk_mg_knn_f32", "nwk_mg_destroy",
)


class NwkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libnwk error {code}: {message}")
        self.code = code


class KnnStats(Structure):
    _fields_ = [
        ("path", c_int32), ("n_fallback_queries", c_int32), ("max_candidates", c_int32),
        ("max_rerank", c_int32), ("total_candidates", c_int64), ("total_rerank", c_int64),
        ("gpu_launches", c_int32), ("max_err_ratio", c_float), ("ms_prep", c_float), ("ms_sample", c_float),
        ("ms_filter", c_float), ("ms_finish", c_float), ("ms_fallback", c_float), ("ms_total", c_float),
        ("n_retry_queries", c_int32), ("reserved_", c_int32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_ if name != "reserved_"}


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("NWK_LIB_PATH") or _build.LIB_PATH     # NWK_LIB_PATH: experiment builds only
    if path == _build.LIB_PATH and (not os.path.exists(path) or (_build._stale() and _can_build())):
        _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing and could not be built; there is no CPU fallback")
    lib = ctypes.CDLL(path)
    lib.nwk_last_error.restype = c_char_p
    lib.nwk_last_error.argtypes = []
    lib.nwk_abi_version.restype = c_int32
    lib.nwk_launch_count.restype = c_int64
    lib.nwk_knn_workspace_bytes.restype = c_size_t
    lib.nwk_knn_workspace_bytes.argtypes = [c_int64, c_int64, c_int32, c_int32, c_int32, c_uint32]
    lib.nwk_knn_f32.restype = c_int32
    lib.nwk_knn_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_uint32, POINTER(KnnStats), c_void_p]
    lib.nwk_merge_workspace_bytes.restype = c_size_t
    lib.nwk_merge_workspace_bytes.argtypes = [c_int32, c_int64, c_int32, c_int32]
    lib.nwk_merge_topk.restype = c_int32
    lib.nwk_merge_topk.argtypes = [c_int32, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]
    lib.nwk_merge_topk_strided.restype = c_int32
    lib.nwk_merge_topk_strided.argtypes = [c_int32, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.nwk_merge_topk_ptrs.restype = c_int32
    lib.nwk_merge_topk_ptrs.argtypes = [c_int32, POINTER(c_void_p), POINTER(c_void_p), c_int64, c_int32, c_int32,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.nwk_topk_state_reset.restype = c_int32
    lib.nwk_topk_state_reset.argtypes = [c_void_p, c_void_p]
    lib.nwk_host_slab_plan.restype = c_int32
    lib.nwk_host_slab_plan.argtypes = [c_int64, c_int64, c_int32, c_int32, c_int64, POINTER(c_int64), POINTER(c_int64),
                                       POINTER(c_int64)]
    lib.nwk_rows_from_columns_f32.restype = c_int32
    lib.nwk_rows_from_columns_f32.argtypes = [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p]
    lib.nwk_xchg_create.restype = c_int32
    lib.nwk_xchg_create.argtypes = [c_int32, c_int32, c_int64, c_int32, POINTER(c_void_p)]
    lib.nwk_xchg_handle_bytes.restype = c_size_t
    lib.nwk_xchg_handle_bytes.argtypes = []
    lib.nwk_xchg_local_handle.restype = c_int32
    lib.nwk_xchg_local_handle.argtypes = [c_void_p, c_void_p]
    lib.nwk_xchg_open_peers.restype = c_int32
    lib.nwk_xchg_open_peers.argtypes = [c_void_p, c_void_p]
    lib.nwk_xchg_next_slot.restype = c_int32
    lib.nwk_xchg_next_slot.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p)]
    lib.nwk_xchg_merge.restype = c_int32
    lib.nwk_xchg_merge.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.nwk_xchg_check.restype = c_int32
    lib.nwk_xchg_check.argtypes = [c_void_p]
    lib.nwk_xchg_destroy.restype = c_int32
    lib.nwk_xchg_destroy.argtypes = [c_void_p]
    lib.nwk_host_alloc.restype = c_int32
    lib.nwk_host_alloc.argtypes = [c_size_t, c_int32, POINTER(c_void_p)]
    lib.nwk_host_free.restype = c_int32
    lib.nwk_host_free.argtypes = [c_void_p]
    lib.nwk_mg_create.restype = c_int32
    lib.nwk_mg_create.argtypes = [POINTER(c_int32), c_int32, POINTER(c_void_p)]
    lib.nwk_mg_knn_f32.restype = c_int32
    lib.nwk_mg_knn_f32.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_int64), c_void_p, c_int64, c_int32, c_int32, c_int32,
                                   c_int64, c_void_p, c_void_p, c_uint32, POINTER(KnnStats)]
    lib.nwk_mg_destroy.restype = c_int32
    lib.nwk_mg_destroy.argtypes = [c_void_p]
    lib.nwk_host_release.restype = c_int32
    lib.nwk_host_release.argtypes = []
    lib.nwk_topk_state_create.restype = c_int32
    lib.nwk_topk_state_create.argtypes = [c_int64, c_int32, POINTER(c_void_p)]
    lib.nwk_topk_state_update.restype = c_int32
    lib.nwk_topk_state_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nwk_maxsim_workspace_bytes.restype = c_size_t
    lib.nwk_maxsim_workspace_bytes.argtypes = [c_int64, c_int32, c_int64, c_int32, c_int32, c_int32]
    lib.nwk_maxsim_f32.restype = c_int32
    lib.nwk_maxsim_f32.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32,
                                   c_int64, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(KnnStats), c_void_p]
    lib.nwk_topk_state_search.restype = c_int32
    lib.nwk_topk_state_search.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int64, c_void_p,
                                          c_size_t, c_uint32, POINTER(KnnStats), c_void_p]
    lib.nwk_topk_state_finalize.restype = c_int32
    lib.nwk_topk_state_finalize.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nwk_topk_state_destroy.restype = c_int32
    lib.nwk_topk_state_destroy.argtypes = [c_void_p]
    lib.nwk_knn_host_f32.restype = c_int32
    lib.nwk_knn_host_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int64,
                                     c_int64, c_void_p, c_void_p, c_uint32, POINTER(KnnStats)]
    lib.nwk_synth_fill.restype = c_int32
    lib.nwk_synth_fill.argtypes = [c_void_p, c_int64, c_int32, c_uint64, c_int64, c_int32, c_void_p]
    lib.nwk_debug_tc_tile.restype = c_int32
    lib.nwk_debug_tc_tile.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int64, c_int64, c_void_p,
                                      c_int32, c_void_p]
    lib.nwk_validate_pairs.restype = c_int32
    lib.nwk_validate_pairs.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32,
                                       ctypes.c_double, ctypes.c_double, POINTER(c_int64), c_void_p]
    if lib.nwk_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.nwk_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def _can_build() -> bool:
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def check(code: int) -> None:
    if code != 0:
        raise NwkError(code, load().nwk_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().nwk_launch_count())
