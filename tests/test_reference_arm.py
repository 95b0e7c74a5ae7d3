"""The reference arm of bench.py (`--impl reference`): the UNMODIFIED reference tree driven through its own
DistributedOptimizer / AllReducer / models with the mpi4py shim, here in the CPU dry-run mode on gloo world 2."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_reference_arm_without_a_gpu_reports_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stderr[-500:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["impl"] == "reference" and "unavailable" in out


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "Ok-Topk", "VGG")), reason="reference tree not installed")
def test_reference_arm_cpu_dry_run_world2_emits_the_canonical_line():
    env = {**os.environ, "OKTOPK_REF_CPU_TEST": "1", "OKTOPK_REF_DENSE_WARMUP": "1", "OKTOPK_BENCH_EXTRA": "0",
           "CUDA_VISIBLE_DEVICES": ""}
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "2", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-800:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["value"] > 0
    # the same config keys as our arm (bench.canonical_config), the losses at the marked steps, the e2e block
    sys.path.insert(0, ROOT)
    import bench
    ours = bench.canonical_config("vgg16", "vgg16", "cifar10", 16, 2, 128, "oktopk", 0.001, 14728266, 1, bench.SPARSE_PHASE)
    assert out["config"] == ours
    assert len(out["loss"]) >= 1 and all(v == v for v in out["loss"].values())
    assert out["e2e"]["h2d_bytes_per_step"] > 0
