"""ctypes binding of include/vers_device.h (libvers_b200.so).

There is no CPU fallback: if the shared library is missing or no CUDA device is present every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvers_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_PANIC, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
METRIC_L2SQ, METRIC_COSINE = 0, 1
MAX_TOPK = 128
SYNTH_UNIFORM, SYNTH_CLUSTERED = 0, 1
REDUCE_CHAINED, REDUCE_ALLREDUCE = 0, 1
KF_LIST_SCAN, KF_FLAT_SCAN, KF_ASSIGN, KF_SUMS, KF_LSH_HASH, KF_PROBE, KF_CAND_SCAN, KF_RERANK = range(8)


class VersError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vers_b200 error {code}: {msg}")
        self.code = code


class VersPanic(VersError):
    """The reference panics on these inputs (unwrap on None / index out of bounds)."""


u64, u32, i32, f32 = C.c_uint64, C.c_uint32, C.c_int32, C.c_float
vp = C.c_void_p
pvp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is int32 unless listed in _RESTYPES)
SIGNATURES = {
    "vers_last_error": [],
    "vers_abi_version": [],
    "vers_ctx_create": [i32, pvp],
    "vers_ctx_destroy": [vp],
    "vers_ctx_set_stream": [vp, vp],
    "vers_ctx_sync": [vp],
    "vers_ctx_launch_count": [vp, C.POINTER(u64)],
    "vers_ctx_enable_timing": [vp, i32],
    "vers_ctx_kernel_ms": [vp, i32, C.POINTER(f32), C.POINTER(u64)],
    "vers_dataset_upload": [vp, vp, u64, u32, u32, u64, pvp],
    "vers_dataset_synth": [vp, u64, u64, u32, u32, u64, u64, u32, i32, pvp],
    "vers_dataset_info": [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)],
    "vers_dataset_normalize": [vp],
    "vers_dataset_download": [vp, u64, u64, vp, u32],
    "vers_dataset_device_ptr": [vp, pvp],
    "vers_dataset_wrap_device": [vp, vp, u64, u32, u64, pvp],
    "vers_dataset_free": [vp],
    "vers_flat_search": [vp, vp, u32, u32, u32, u32, vp, vp, vp],
    "vers_flat_search_dev": [vp, vp, u32, u32, u32, vp, vp, vp],
    "vers_flat_set_mode": [vp, i32],
    "vers_debug_peer_times": [vp, vp],
    "vers_pair_distances_simd": [vp, vp, u32, u32, vp, vp, u64, u32, vp],
    "vers_pair_distances_simd_dev": [vp, vp, u32, u32, vp, vp, u64, u32, vp, vp],
    "vers_flat_last_search_stats": [vp, vp],
    "vers_kmeans_create": [vp, u32, pvp],
    "vers_kmeans_free": [vp],
    "vers_kmeans_init_from_rows": [vp, vp],
    "vers_kmeans_set_centroids": [vp, vp, u32],
    "vers_kmeans_get_centroids": [vp, vp, u32],
    "vers_kmeans_get_assignments": [vp, vp],
    "vers_kmeans_centroids_device_ptr": [vp, pvp, C.POINTER(u32)],
    "vers_kmeans_assign_device_ptr": [vp, pvp],
    "vers_kmeans_assign_step": [vp],
    "vers_kmeans_set_mode": [vp, i32],
    "vers_kmeans_last_assign_stats": [vp, C.POINTER(u64)],
    "vers_kmeans_sums_step_dev": [vp, vp, vp],
    "vers_kmeans_finalize_step_dev": [vp, vp, vp, C.POINTER(u32)],
    "vers_kmeans_cost_step": [vp, C.POINTER(f32)],
    "vers_kmeans_fit": [vp, u32, C.POINTER(u32)],
    "vers_kmeans_assign": [vp, vp, u32, u32, vp],
    "vers_kmeans_update": [vp, vp, u32, vp, vp],
    "vers_ivf_build_index": [vp, u32, u32, u32, vp, pvp],
    "vers_ivf_from_kmeans": [vp, pvp],
    "vers_ivf_from_parts": [vp, vp, u32, u32, vp, pvp],
    "vers_ivf_free": [vp],
    "vers_ivf_info": [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u32), C.POINTER(f32), C.POINTER(u32)],
    "vers_ivf_get_centroids": [vp, vp, u32],
    "vers_ivf_get_assignments": [vp, vp],
    "vers_ivf_get_list_sizes": [vp, vp],
    "vers_ivf_get_list": [vp, u32, vp, vp, u32],
    "vers_ivf_last_search_stats": [vp, vp],
    "vers_ivf_from_parts_dev": [vp, vp, u32, vp, vp, pvp],
    "vers_ivf_set_mode": [vp, i32],
    "vers_ivf_search": [vp, vp, u32, u32, u32, u32, vp, vp, vp],
    "vers_ivf_search_dev": [vp, vp, u32, u32, u32, vp, vp, vp],
    "vers_ivf_probe_dev": [vp, vp, u32, u32, vp],
    "vers_ivf_search_probed_dev": [vp, vp, u32, u32, u32, vp, vp, vp, vp],
    "vers_ivf_add": [vp, vp, u64, C.POINTER(u64), C.POINTER(u32)],
    "vers_ivf_add_batch": [vp, vp, u64, u32, vp, vp],
    "vers_topk_merge_dev": [vp, vp, vp, u32, u64, u64, u32, u32, vp, vp, vp],
    "vers_peer_create": [vp, u32, u32, u64, pvp, vp],
    "vers_peer_connect": [vp, vp],
    "vers_peer_gather_merge_dev": [vp, vp, vp, u32, u32, vp, vp, vp],
    "vers_peer_free": [vp],
    "vers_comm_unique_id": [vp],
    "vers_comm_create": [vp, u32, u32, vp, pvp],
    "vers_comm_destroy": [vp],
    "vers_comm_info": [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_double)],
    "vers_comm_barrier": [vp],
    "vers_comm_max_f64": [vp, C.POINTER(C.c_double)],
    "vers_sharded_kmeans_fit": [vp, vp, vp, u32, i32, C.POINTER(u32)],
    "vers_sharded_kmeans_cost": [vp, vp, C.POINTER(f32)],
    "vers_sharded_ivf_build": [vp, vp, pvp],
    "vers_sharded_list_owners": [vp, u32, u32, vp],
    "vers_sharded_ivf_search": [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp],
    "vers_sharded_ivf_search_dev": [vp, vp, vp, u32, u32, u32, vp, vp, vp],
    "vers_sharded_lsh_search": [vp, vp, vp, u32, u32, u32, vp, vp, vp],
    "vers_lsh_hash": [vp, vp, u32, u32, vp, vp],
    "vers_lsh_hash_dev": [vp, vp, u32, vp, vp],
    "vers_lsh_build_index": [vp, vp, u64, u32, u32, vp, u32, u32, u64, pvp],
    "vers_lsh_free": [vp],
    "vers_lsh_info": [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u64)],
    "vers_lsh_flatten": [vp, u32, vp, vp, vp, vp, vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)],
    "vers_lsh_search": [vp, vp, u32, u32, u32, vp, vp, vp],
    "vers_lsh_add": [vp, vp, u64],
    "vers_lsh_get_values": [vp, vp, u32, vp],
    "vers_lsh_from_parts": [vp, vp, u64, u32, u32, vp, u32, u32, u64, vp, vp, vp, vp, vp, vp, pvp],
}
_RESTYPES = {"vers_last_error": C.c_char_p}

_lib = None


def lib() -> C.CDLL:
    """Load libvers_b200.so (built in-tree by __graft_entry__.build() / vers_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(vers_b200 has no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, i32)
    _lib = L
    return L


def check(rc: int):
    if rc == OK:
        return
    msg = lib().vers_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == ERR_PANIC:
        raise VersPanic(rc, msg)
    raise VersError(rc, msg)


def ptr(a) -> C.c_void_p:
    """numpy array -> void*"""
    return a.ctypes.data_as(C.c_void_p)
