"""ptoy_b200 — B200-native (sm_100a) engine for ptoy's particle-stepping hot path.

``Particles`` mirrors the reference ``class Particles`` (src/particles.h) over the C ABI of
``libptoy_b200.so`` (include/ptoy_b200.h); ``scenes`` generates the synthetic workloads of
BASELINE.json.  The package holds only what the hot path needs.
"""
from .particles import (Particles, PtoyError, StripGroup, load_library, LIB_PATH,  # noqa: F401
                        nccl_unique_id, partition_columns)
from . import scenes  # noqa: F401

__all__ = ["Particles", "PtoyError", "StripGroup", "load_library", "LIB_PATH", "scenes",
           "nccl_unique_id", "partition_columns"]
