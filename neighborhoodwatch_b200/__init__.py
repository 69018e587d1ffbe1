"""neighborhoodwatch_b200 -- B200-native exact brute-force kNN for neighborhoodwatch's
ground-truth hot path (the fused distance + top-k behind cu_knn.py / cu_knn_ds.py and the
partial-result merge of merge.py), hand-written sm_100a CUDA behind a C ABI (include/nwk.h).

Importing the package is cheap and works without a GPU; every compute entry point needs
the built ``libnwk.so`` and a B200 and fails loudly otherwise -- there is no CPU fallback.
"""
from ._native import NwkError, lib_path  # noqa: F401

__all__ = ["NwkError", "lib_path"]
__version__ = "0.1.0"
