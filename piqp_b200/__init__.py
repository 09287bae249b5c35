"""piqp_b200 -- B200 (sm_100a) KKT factorise-and-solve backend for PIQP.

Product code: the CUDA library (csrc/ -> libpiqp_b200.so, C-ABI in include/piqp_b200.h) and the thin
host-side mirrors of the reference interfaces in this package.  No CPU fallback exists.
"""
from ._lib import Info, Settings, Stats, build, lib  # noqa: F401
from .backend import sparse_ldlt_symbolic, LDLTNoPivot, DenseKKT, MultistageKKT, SparseKKT, KKTSolverBase, KKT_UPDATE_A, KKT_UPDATE_G, KKT_UPDATE_NONE, KKT_UPDATE_P  # noqa: F401
from .solver import (DenseSolverBatched, SparseSolverBatched, PIQP_DUAL_INFEASIBLE, PIQP_MAX_ITER_REACHED, PIQP_NUMERICS,  # noqa: F401
                     PIQP_PRIMAL_INFEASIBLE, PIQP_SOLVED, PIQP_UNSOLVED)
