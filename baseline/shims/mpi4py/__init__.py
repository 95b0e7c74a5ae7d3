"""Minimal ``mpi4py`` stand-in for running the UNMODIFIED reference (Shigangli/Ok-Topk) on a box without MPI.

The image has neither an MPI library nor the mpi4py wheel (no network), so the reference's
``from mpi4py import MPI`` cannot be satisfied by the real package.  This shim maps the handful of
``MPI.COMM_WORLD`` calls the reference makes on *host NumPy buffers*  onto
``torch.distributed`` with the gloo backend on CPU tensors that alias those buffers -- i.e. the same
host-staged communication pattern, just with gloo instead of Cray-MPICH underneath.  It is NOT part of
the product; it only exists so that the baseline arm of bench.py can execute the reference's own
``distributed_optimizer.py`` / ``allreducer.py`` / ``compression.py`` byte for byte.
"""
from . import MPI  # noqa: F401

__all__ = ["MPI"]
