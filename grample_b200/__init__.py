"""grample_b200 — B200-native Gibbs sweep hot path of CraigKelly/grample behind a C ABI.

`core` wraps the C ABI one call per method; `sampler` mirrors the reference's `sampler`
package API (NewGibbsSimple, NewChain, MergeChains, ...) on top of it.
"""
from ._lib import (CHAINS_HISTORY, F32, F64, HYBRID, TABLE, TABLE_BITS, HELLINGER, JS, MAX_ABS, MAX_CARD, MEAN_ABS, NEIGHBOR_VAR_MAX,
                   GrampleError, exported_symbols)
from .core import Chains, Comm, Fleet, Model, device_count, error_suite, mar_load
from .ising import ising_torus

__all__ = ["Chains", "Comm", "Fleet", "Model", "device_count", "error_suite", "mar_load", "ising_torus", "GrampleError",
           "exported_symbols", "F32", "F64", "TABLE", "HYBRID", "TABLE_BITS", "HELLINGER", "JS", "MAX_ABS", "MEAN_ABS", "CHAINS_HISTORY",
           "NEIGHBOR_VAR_MAX", "MAX_CARD"]
