"""gams_b200 — B200 (sm_100a) drop-in for the data-parallel hot path of wang-q/gams.

The product is gams_b200/libgams_cuda.so (C ABI in include/gams_cuda.h).  This package is the
thin Python host mirror used by tests and bench.py: the same names and argument meaning as the
reference's `gams::` functions for this path (sliding, thresholding_algo, count_rg, center_sw ...),
each calling the CUDA library.  There is no CPU fallback: importing `gams_b200.api` without the
built library, or calling it without a B200, raises.
"""
from ._lib import GamsCudaError, lib_path, load  # noqa: F401

__version__ = "0.1.0"
