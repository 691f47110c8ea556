"""searchlite_b200 — B200-native engine for searchlite's BM25 top-k hot path.

The product is ``libsearchlite_gpu.so`` (C ABI in ``include/searchlite_gpu.h``, CUDA in
``searchlite_b200/csrc``).  This package is the thin Python host layer used by tests and
``bench.py``: a ctypes binding (`engine`), the synthetic corpus / query generators
(`synth`) and the multi-GPU shard driver (`shard`).  Nothing here imports ``oracle/``.
"""
from .engine import (  # noqa: F401
    EXECUTION,
    GpuIndex,
    QueryBatch,
    SearchliteGpuError,
    SegmentData,
    load_library,
)

__all__ = ["EXECUTION", "GpuIndex", "QueryBatch", "SearchliteGpuError", "SegmentData", "load_library"]
