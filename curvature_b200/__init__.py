"""curvature_b200 -- B200-native (sm_100a) implementation of the Fisher-estimation hot path of
DLR-RM/curvature behind the reference's Python API.  Importing this package loads the CUDA shared
library; it raises if the library has not been built (there is no fallback implementation)."""
from . import _native
from .curvatures import Curvature, Diagonal, KFAC, EFB, INF, FactorArena
from .utils import get_eigenvectors, get_eigenvalues, eigendecompose, kron
from .parallel import allreduce_arena, shard_indices, invert_plan, invert_plan_two_rounds, allgather_segments
from .io import save_factors, load_factors
from .evaluate import eval_nn, eval_bnn

__all__ = ["Curvature", "Diagonal", "KFAC", "EFB", "INF", "FactorArena", "get_eigenvectors", "get_eigenvalues",
           "eigendecompose", "kron", "allreduce_arena", "shard_indices", "invert_plan", "invert_plan_two_rounds", "allgather_segments", "save_factors", "load_factors", "eval_nn", "eval_bnn"]
