"""vers_b200 — B200-native (sm_100a) device layer behind vers' ``Index<N>`` trait.

The product is the C-ABI shared library ``libvers_b200.so`` (include/vers_device.h); this package is the thin
host-side mirror of the reference's interface used by the tests, bench.py and Python callers.  There is no CPU
fallback: importing works without a GPU (so the CPU test tier can check the ABI), every compute call needs one.
"""
from ._abi import (ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_PANIC, ERR_UNSUPPORTED, LIB_PATH, MAX_TOPK, METRIC_COSINE,  # noqa: F401
                   METRIC_L2SQ, SYNTH_CLUSTERED, SYNTH_UNIFORM, VersError, VersPanic, lib)
from .index import (ANNIndex, Context, Dataset, IVFFlatIndex, KMeans, assign_to_clusters, default_context,  # noqa: F401
                    lsh_hash, pair_distances_simd, search_exhaustive, search_exhaustive_batch, synth_init_rows, update_centroids)

__all__ = [
    "ANNIndex", "Context", "Dataset", "IVFFlatIndex", "KMeans", "VersError", "VersPanic", "assign_to_clusters",
    "default_context", "lib", "lsh_hash", "pair_distances_simd", "search_exhaustive", "search_exhaustive_batch", "synth_init_rows",
    "update_centroids",
]
