"""Install the reference (Shigangli/Ok-Topk) under ``baseline/_ref`` for the ``--impl reference`` arm.

1. The prescribed offline ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
   baseline/_ref /root/reference`` is attempted first.  It fails by construction: the reference ships no
   ``setup.py`` / ``pyproject.toml`` ("Directory '/root/reference' is not installable"): it is three
   directories of scripts run in place (``srun python -m mpi4py main_trainer.py``).
2. Fallback = what "installing" such a project means: copy the unmodified source tree to
   ``baseline/_ref/Ok-Topk`` (git-ignored, travels with gpurun).  Its one missing hard dependency, mpi4py
   (no MPI library and no wheel in this image), is satisfied by ``baseline/shims/mpi4py`` at run time.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("OKTOPK_REFERENCE", "/root/reference")


def install(force: bool = False) -> str:
    tree = os.path.join(DST, "Ok-Topk")
    if os.path.isdir(os.path.join(tree, "VGG")) and not force:
        return tree
    os.makedirs(DST, exist_ok=True)
    log = []
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
           "/opt/wheelhouse", "--target", DST, SRC]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        log.append("$ " + " ".join(cmd))
        log.append((r.stdout + r.stderr).strip()[-2000:])
        log.append("exit code %d" % r.returncode)
    except Exception as e:  # noqa: BLE001
        log.append("pip attempt raised %r" % (e,))
    if not os.path.isdir(SRC):
        raise RuntimeError("reference source %s not found" % SRC)
    if os.path.isdir(tree):
        shutil.rmtree(tree)
    shutil.copytree(SRC, tree, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    log.append("fallback: copied the unmodified tree %s -> %s" % (SRC, tree))
    with open(os.path.join(DST, "INSTALL_LOG.txt"), "w") as f:
        f.write("\n".join(log) + "\n")
    return tree


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
