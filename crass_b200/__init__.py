"""crass_b200 -- B200-native read-scanning hot path of crass (direct-repeat search + singleton scan).

The product is the CUDA library ``libcrass_b200.so`` (sources in ``crass_b200/csrc``, C-ABI in
``include/crass_b200.h``).  This package is its ctypes binding plus the seeded synthetic read
generator and the multi-GPU driver used by the benchmarks.
"""
from .api import (Automaton, Batch, Context, CrassB200Error, Engine, Hit, HIT_DTYPE, Params, Results, device_count, lib,  # noqa: F401
                  non_redundant_set, pack_reads, run_files_multi)

__all__ = ["Automaton", "Batch", "Context", "CrassB200Error", "Engine", "Hit", "HIT_DTYPE", "Params", "Results", "device_count", "lib",
           "non_redundant_set", "pack_reads", "run_files_multi"]
