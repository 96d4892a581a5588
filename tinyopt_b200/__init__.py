"""tinyopt_b200 — B200-native batched dense-NLLS Levenberg-Marquardt inner loop.

Host-side mirror of the reference interface for this path (tinyopt::Options / Output / StopReason,
Optimize -> optimize_batch, SolverLM -> BatchSolver) over the C-ABI in include/tinyopt_b200.h.
PyTorch is used for device memory and streams only.
"""
from .api import (  # noqa: F401
    BatchSolver, Context, Options, Output, StopReason, TILE32, PROBLEM_MAJOR, TinyoptB200Error,
    from_tile32, options, to_tile32,
)
