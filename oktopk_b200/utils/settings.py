"""Global debugging / profiling switches (``VGG/settings.py:6-26``), settable from the environment."""
import os


def _flag(name: str, default: bool = False) -> bool:
    return os.environ.get("OKTOPK_" + name, "1" if default else "0") == "1"


# Only flags that something reads live here (the reference's DEBUG/SPARSE/WARMUP/DELAY_COMM are dead there too).
PREFIX = os.environ.get("OKTOPK_PREFIX", "")   # directory of the PROFILING_GRAD / PROFILING_NORM dumps
TENSORBOARD = _flag("TENSORBOARD")        # utils/metrics.MetricsWriter also writes TensorBoard event files
PROFILING = _flag("PROFILING")            # per-call selected counts / thresholds + 50-call phase means (parallel/allreducer.py)
PROFILING_NORM = _flag("PROFILING_NORM")  # relative error of the sparse result vs true dense top-k + the gtopk/randk norm tuples
PROFILING_GRAD = _flag("PROFILING_GRAD")  # dump local gradient / threshold snapshots at chosen iterations
