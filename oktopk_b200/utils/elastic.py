"""Failure handling around the training loop.

The reference carries three vestigial mechanisms: ``MPI.ERRORS_RETURN`` with nothing catching the error, an
``err_callback(new_num_workers, new_rank)`` that is stored and never invoked (``VGG/allreducer.py:182-183,237``) with
``DLTrainer.update_nworker`` behind it (``VGG/dl_trainer.py:472-493``), and SLURM signal handlers / interrupted-state
save / ``scontrol requeue`` helpers that are never installed (``BERT/bert/main_bert.py:51-153``).  Here they are live:

* **detection** happens on the device: every cross-GPU wait in the kernels is bounded (``OkTopkConfig.peer_timeout_s``),
  a timeout leaves a fault code in the bucket state, ``optimizer.check_faults()`` surfaces it as ``PeerTimeoutError``
  or routes it to the ``err_handler``;
* **signals**: ``install_signal_handlers(trainer, path)`` saves an *interrupted-state* checkpoint (model, optimizer,
  residuals, thresholds, region edges, counters) on SIGUSR1/SIGTERM and optionally requeues the SLURM job;
* **resume / shrink**: ``resume_if_interrupted`` reloads that state; a checkpoint written at world size P loads at
  P' (regions fall back to uniform, ``SparseState.load_state_dict``) and ``Trainer.update_nworker`` re-shards the data.
"""
from __future__ import annotations

import os
import signal
import subprocess
import sys
from typing import Callable, Optional

INTERRUPTED_NAME = "interrupted_state.pth"


def interrupted_path(directory: str, rank: int = 0) -> str:
    return os.path.join(directory, "rank%d_%s" % (rank, INTERRUPTED_NAME))


def save_interrupted_state(trainer, directory: str) -> str:
    os.makedirs(directory, exist_ok=True)
    path = interrupted_path(directory, trainer.rank)
    trainer.save_checkpoint(path, collective=False)        # per-rank file, callable from a signal handler
    return path


def resume_if_interrupted(trainer, directory: str, remove: bool = True) -> bool:
    path = interrupted_path(directory, trainer.rank)
    if not os.path.exists(path):
        return False
    trainer.load_checkpoint(path)
    if remove:
        os.remove(path)
    return True


def requeue_job() -> bool:
    """``scontrol requeue $SLURM_JOB_ID`` when running under SLURM (``main_bert.py:120-153``)."""
    job = os.environ.get("SLURM_JOB_ID")
    if not job:
        return False
    try:
        subprocess.run(["scontrol", "requeue", job], check=True, timeout=30)
        return True
    except Exception:  # noqa: BLE001
        return False


def install_signal_handlers(trainer, directory: str, requeue: bool = False, exit_after: bool = True,
                            on_signal: Optional[Callable[[int], None]] = None):
    """SIGUSR1 (SLURM's pre-timeout notice) and SIGTERM => interrupted-state checkpoint (+ requeue on rank 0)."""

    def handler(signum, _frame):
        path = save_interrupted_state(trainer, directory)
        trainer.logger.warning("signal %d: interrupted state saved to %s", signum, path)
        if on_signal is not None:
            on_signal(signum)
        if requeue and trainer.rank == 0:
            requeue_job()
        if exit_after:
            sys.exit(0)

    prev = {}
    for sig in (signal.SIGUSR1, signal.SIGTERM):
        prev[sig] = signal.signal(sig, handler)
    return prev


def shrink_world(trainer, new_num_workers: int, new_rank: int) -> None:
    """The reference's ``_error_handler`` -> ``update_nworker`` path (``VGG/main_trainer.py:42-44``)."""
    trainer.update_nworker(new_num_workers, new_rank)
