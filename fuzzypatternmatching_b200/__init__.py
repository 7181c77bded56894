"""fuzzypatternmatching_b200 — B200-native LCC/NLCC pattern-matching pruning engine.

Drop-in for the pruning path of HavoqGT's run_pattern_matching_beta: the compute
lives in libpmgpu.so (hand-written sm_100a CUDA behind the C ABI of
include/pmgpu.h); this package is the thin host-side mirror of the reference
driver's interface.  Importing `engine` requires the built library.
"""
from . import patterns  # noqa: F401

__all__ = ["patterns"]
