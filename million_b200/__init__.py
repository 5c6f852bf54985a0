"""million_b200 — B200-native (sm_100a) implementation of MILLION's product-quantized KV-cache hot path.

Mirrors the reference's PyTorch-facing surface (Zhaohui-Xu/MILLION):
    million_b200.pq_utils        <-> scripts/utils/pq_utils.py        (DynamicPQCache, KernelRegistry, sa_*)
    million_b200.paged_pq_utils  <-> scripts/utils/paged_pq_utils.py  (PagedPQCache) + PageManager
    million_b200.bindings        <-> scripts/modeldb/bindings (pybind module `bindings`, named kernels)
over a thin C ABI (include/million_b200.h, libmillion_b200.so).  No CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the .so until first use)
