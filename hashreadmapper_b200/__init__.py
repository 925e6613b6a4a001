"""hashreadmapper_b200 -- B200-native (sm_100a) implementation of hashreadmapper's read-mapping hot path.

The product is libhrm_b200.so (hand-written CUDA kernels behind the C ABI of include/hrm_b200.h);
this package is the thin host-side mirror of the reference's handle API plus the multi-GPU plumbing.
Importing `hashreadmapper_b200.api` requires the built library: there is no CPU / PyTorch fallback.
"""
from ._lib import (HrmError, load, check, CONV_NONE, CONV_CT, CONV_GA, ORIENT_FORWARD, ORIENT_REVCOMP, ORIENT_NONE,
                   MAPPER_SW, MAPPER_EDLIB, MAPPED_DTYPE, ALIGN_DTYPE, RECORD_DTYPE, SIGNATURES)

__all__ = ["HrmError", "load", "check", "CONV_NONE", "CONV_CT", "CONV_GA", "ORIENT_FORWARD", "ORIENT_REVCOMP",
           "ORIENT_NONE", "MAPPER_SW", "MAPPER_EDLIB", "MAPPED_DTYPE", "ALIGN_DTYPE", "RECORD_DTYPE", "SIGNATURES"]
