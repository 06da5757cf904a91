"""physher_b200 -- B200-native (sm_100a) tree-likelihood path behind physher's SingleTreeLikelihood API.

The product is the C-ABI shared library `libphysher_b200.so` (include/physher_b200.h); this package
holds its sources (csrc/), the in-tree build and a ctypes mirror of the reference interface used by
the tests and the benchmark.  There is no CPU fallback: importing works anywhere, computing needs a GPU.
"""
from .treelikelihood import (  # noqa: F401
    Comm,
    FLAG_TREE_MODEL,
    KERNELS_AUTO,
    KERNELS_FUSED,
    KERNELS_GENERIC,
    OPT_INCREMENTAL,
    PhysherB200Error,
    SingleTreeLikelihood,
    TreeLikelihoodGroup,
    compress_patterns,
    device_count,
    load_library,
    nccl_version,
)

__all__ = [
    "SingleTreeLikelihood", "TreeLikelihoodGroup", "Comm", "nccl_version", "PhysherB200Error", "load_library", "device_count", "compress_patterns",
    "FLAG_TREE_MODEL", "KERNELS_AUTO", "KERNELS_GENERIC", "KERNELS_FUSED", "OPT_INCREMENTAL",
]
