"""Global debugging / profiling switches (``VGG/settings.py:6-26``), settable from the environment."""
import os


def _flag(name: str, default: bool = False) -> bool:
    return os.environ.get("OKTOPK_" + name, "1" if default else "0") == "1"


DEBUG = _flag("DEBUG")
SPARSE = _flag("SPARSE")
WARMUP = _flag("WARMUP", True)
DELAY_COMM = 1
PREFIX = os.environ.get("OKTOPK_PREFIX", "")
TENSORBOARD = _flag("TENSORBOARD")
PROFILING = _flag("PROFILING")            # per-iteration selected counts / thresholds
PROFILING_NORM = _flag("PROFILING_NORM")  # relative error of the sparse result vs true dense top-k
PROFILING_GRAD = _flag("PROFILING_GRAD")  # dump raw gradient / threshold snapshots
