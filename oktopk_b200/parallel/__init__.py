"""Data-parallel sparse-allreduce machinery: world/bootstrap, per-bucket state, the schemes on
library collectives, the single-process oracle, symmetric peer memory and the fused CUDA engine."""
from .state import SparseState, uniform_boundaries, offsets_of  # noqa: F401
from .world import World, init, world, rank, size, shutdown  # noqa: F401
