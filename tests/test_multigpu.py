"""Multi-GPU tests (>= 2 devices): the fused peer-memory kernels across real NVLink peers versus the
single-process oracle.  At P=2 the sparse reduction is bitwise reproducible (a+b == b+a), so results,
residuals, thresholds and region boundaries must match the oracle exactly."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mp_util import run_distributed  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _grad(it, rank, n):
    g = torch.Generator().manual_seed(1000 * it + rank)
    x = torch.randn(n, generator=g)
    # non-uniform magnitude profile so that balanced regions differ from uniform ones
    ramp = torch.linspace(0.2, 2.0, n)
    return x * ramp


def _engine_worker(rank, P, name, n, iters, cfg_kw):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine
    from oktopk_b200.parallel.world import World
    w = World()
    cfg = OkTopkConfig(**cfg_kw)
    eng = CudaBucketEngine(n, cfg, w, name="t")
    outs, stats = [], []
    for it in range(iters):
        eng.grad.copy_(_grad(it, rank, n).cuda())
        torch.cuda.synchronize()
        w.barrier()
        eng.reduce(name)
        torch.cuda.synchronize()
        outs.append(eng.grad.cpu().clone())
        stats.append(eng.stats())
    res = eng.residual.cpu().clone()
    w.barrier()
    eng.close()
    return outs, res, stats


def _check(name, P, n, iters, cfg_kw, exact=True):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.oracle import run_oracle
    from oktopk_b200.parallel.state import SparseState
    got = run_distributed(_engine_worker, P, (name, n, iters, cfg_kw), backend="nccl", timeout=600)
    cfg = OkTopkConfig(**cfg_kw)
    states = [SparseState(n, P) for _ in range(P)]
    for it in range(iters):
        ref = run_oracle(name, [_grad(it, r, n) for r in range(P)], states, cfg)
        for r in range(P):
            out = got[r][0][it]
            if exact:
                bad = int((out != ref[r]).sum())
                assert bad == 0, "%s it %d rank %d: %d mismatches, stats %s" % (name, it, r, bad, got[r][2][it])
                assert got[r][2][it]["edges"] == states[r].region_offsets + [n], (it, got[r][2][it]["edges"], states[r].region_offsets)
            else:
                # >2 contributions: the summation order differs from the oracle's, so an entry within an ulp of a
                # threshold may flip -- tolerate a handful of such borderline elements, nothing else
                bad = int((~torch.isclose(out, ref[r], rtol=1e-5, atol=1e-6)).sum())
                assert bad <= 4, "%s it %d rank %d: %d elements differ from the oracle" % (name, it, r, bad)
            assert got[r][2][it]["overflow_send"] == 0 and got[r][2][it]["overflow_gather"] == 0
    if exact and states[0].residual is not None:
        for r in range(P):
            assert int((got[r][1] != states[r].residual).sum()) == 0, "residual mismatch rank %d" % r
    return got


@pytest.mark.parametrize("mode", ["list", "scan"])
def test_oktopk_two_gpus_global_selection_paths(mode):
    _check("oktopk", 2, 700_001, 7, dict(density=0.004, local_recompute_interval=3, global_recompute_interval=4,
                                         repartition_interval=4, gselect_mode=mode, slot_factor=64, gather_factor=64))


@pytest.mark.parametrize("P", [4, 8])
def test_oktopk_many_gpus_matches_oracle(P):
    """P = 4 / 8: indices must match the oracle exactly; values up to the summation order of 4-8 contributions."""
    if torch.cuda.device_count() < P:
        pytest.skip("needs %d GPUs" % P)
    kw = dict(density=0.01, local_recompute_interval=4, global_recompute_interval=4, repartition_interval=4,
              slot_factor=64, gather_factor=64)
    _check("oktopk", P, 1_000_003, 9, kw, exact=False)


@pytest.mark.parametrize("name", ["topkSA", "topkA", "gaussiank", "none"])
def test_baselines_four_gpus(name):
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    kw = dict(density=0.01, slot_factor=64, gather_factor=64)
    if name == "gaussiank":
        got = run_distributed(_engine_worker, 4, (name, 400_000, 3, kw), backend="nccl", timeout=600)
        for it in range(3):
            for r in range(1, 4):
                assert torch.equal(got[r][0][it], got[0][0][it])
        return
    _check(name, 4, 400_000, 4, kw, exact=False)


@pytest.mark.parametrize("pull", ["tma", "ldg"])
def test_oktopk_two_gpus_matches_oracle(pull):
    kw = dict(density=0.01, local_recompute_interval=4, global_recompute_interval=4, repartition_interval=4, pull_mode=pull,
              slot_factor=64, gather_factor=64)
    got = _check("oktopk", 2, 1_000_000, 10, kw)
    # communication volume of the threshold-reuse iterations stays under the 6k(P-1)/P bound (+ slack for k drift)
    k = 10_000
    for it in (0, 4, 8):
        assert got[0][2][it]["volume_elems"] <= 2 * 6 * k, got[0][2][it]


def test_oktopk_two_gpus_unfused_and_deterministic():
    kw = dict(density=0.02, local_recompute_interval=3, global_recompute_interval=5, repartition_interval=2,
              fused=False, deterministic=True, slot_factor=64, gather_factor=64)
    _check("oktopk", 2, 300_000, 7, kw)


@pytest.mark.parametrize("name", ["topkSA", "gaussiankSA", "topkA", "topkAopt", "topkA2", "gtopk", "none"])
def test_baseline_schemes_two_gpus_match_oracle(name):
    kw = dict(density=0.01, topkaopt_recompute_interval=3, slot_factor=64, gather_factor=64)
    _check(name, 2, 400_000, 4, kw)


def test_gaussiank_two_gpus_close():
    from oktopk_b200.config import OkTopkConfig
    got = run_distributed(_engine_worker, 2, ("gaussiank", 400_000, 3, dict(density=0.01, slot_factor=64, gather_factor=64)), backend="nccl", timeout=600)
    # both ranks must hold the identical result
    for it in range(3):
        assert torch.equal(got[0][0][it], got[1][0][it])
        nnz = int((got[0][0][it] != 0).sum())
        assert 0.5 * 4000 < nnz < 3 * 2 * 4000, nnz


def _opt_worker(rank, P, compressor, steps):
    import oktopk_b200 as okt
    from oktopk_b200.models import create_net
    okt.init()
    torch.manual_seed(0)
    net, _ = create_net(10, "resnet20")
    net = net.cuda()
    okt.broadcast_parameters(net)
    cfg = okt.preset("vgg16", density=0.02, warmup_iters=2, bucket_elems=100_000)     # several buckets -> overlap path
    opt = okt.DistributedOptimizer(torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4),
                                   named_parameters=net.named_parameters(), compression=okt.compressors[compressor],
                                   is_sparse=compressor != "none", cfg=cfg)
    torch.manual_seed(100 + rank)
    x = torch.randn(16, 3, 32, 32, device="cuda"); y = torch.randint(0, 10, (16,), device="cuda")
    losses = []
    for it in range(steps):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    chk = float(sum(p.double().sum() for p in net.parameters()))
    nb = len(opt._buckets)
    opt.close()
    return losses, chk, nb


@pytest.mark.parametrize("compressor", ["oktopk", "none"])
def test_distributed_optimizer_two_gpus_replicas_stay_identical(compressor):
    r = run_distributed(_opt_worker, 2, (compressor, 12), backend="nccl", timeout=600)
    assert r[0][2] > 1
    assert r[0][1] == r[1][1], "replicas diverged: %r vs %r" % (r[0][1], r[1][1])
    assert r[0][0][-1] < r[0][0][0]


def _fault_worker(rank, P):
    """Fault injection: rank 1 'dies' (never enters the second reduction).  Rank 0 must not hang: its bounded wait
    records a fault code and ``check_fault`` raises."""
    import time
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine, PeerTimeoutError
    from oktopk_b200.parallel.world import World
    w = World()
    cfg = OkTopkConfig(density=0.01, peer_timeout_s=1.5, slot_factor=64, gather_factor=64)
    eng = CudaBucketEngine(200_000, cfg, w, name="t")
    eng.grad.copy_(_grad(0, rank, 200_000).cuda())
    eng.reduce("oktopk")
    torch.cuda.synchronize()
    eng.check_fault()                                    # healthy round: no fault
    w.barrier()
    outcome = "skipped"
    if rank == 0:
        eng.grad.copy_(_grad(1, rank, 200_000).cuda())
        t0 = time.time()
        eng.reduce("oktopk")
        torch.cuda.synchronize()                         # returns: the kernel gave up waiting instead of hanging
        dt = time.time() - t0
        try:
            eng.check_fault()
            outcome = "no-fault"
        except PeerTimeoutError as e:
            outcome = "fault:%s:%.1f" % (str(e)[:40], dt)
        eng.clear_fault()
        torch.cuda.synchronize()
        assert eng.stats()["fault"] == 0
    else:
        time.sleep(4.0)
    w.barrier()
    eng.close()
    return outcome


def test_peer_timeout_is_detected_not_hung():
    got = run_distributed(_fault_worker, 2, (), backend="nccl", timeout=120)
    assert got[0].startswith("fault:"), got
    assert float(got[0].rsplit(":", 1)[1]) < 20.0
    assert got[1] == "skipped"


# ======================================================================================================
# round 2: native gTopk / TopkA2, lossless exchange + conservation across ranks, dense fallback, stress
# ======================================================================================================
@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("name", ["gtopk", "topkA2"])
def test_native_tree_and_reselect_schemes_match_oracle(name, P):
    if torch.cuda.device_count() < P:
        pytest.skip("needs %d GPUs" % P)
    _check(name, P, 400_000, 4, dict(density=0.01), exact=(P == 2))


def test_topkdsa_dense_fallback():
    """Density 0.4: the reduced regions hold >= n/3 non-zeros, the kernel's final phase takes the dense-allgather path
    (reference VGG/allreducer.py:1311-1353) -- same numbers as the oracle, `dense_fallback` reported."""
    got = _check("topkDSA", 2, 300_000, 3, dict(density=0.4, compressor="topkDSA", dense_switch_density=0.0))
    assert all(got[r][2][it]["dense_fallback"] == 1 for r in range(2) for it in range(3)), got[0][2]
    got = _check("topkDSA", 2, 300_000, 2, dict(density=0.01, compressor="topkDSA"))
    assert all(got[r][2][it]["dense_fallback"] == 0 for r in range(2) for it in range(2))


def _conservation_worker(rank, P, n, iters, cfg_kw):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine
    from oktopk_b200.parallel.world import World
    w = World()
    eng = CudaBucketEngine(n, OkTopkConfig(**cfg_kw), w, name="t")
    accs, outs, ress, stats = [], [], [], []
    res_prev = torch.zeros(n, device="cuda")
    for it in range(iters):
        scale = 1.0 if it < 3 else (10.0 if it < 6 else 100.0)
        x = (_grad(it, rank, n) * scale).cuda()
        accs.append((x + res_prev).double().cpu())
        eng.grad.copy_(x)
        torch.cuda.synchronize()
        w.barrier()
        eng.reduce("oktopk")
        torch.cuda.synchronize()
        outs.append(eng.grad.cpu().clone())
        ress.append(eng.residual.double().cpu().clone())
        stats.append(eng.stats())
        res_prev = eng.residual.clone()
    w.barrier()
    eng.close()
    return accs, outs, ress, stats


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("slot_factor", [0.0, 1.0])
def test_mass_is_conserved_across_ranks_under_scale_jumps(P, slot_factor):
    """sum_r acc_r == sum_r residual_r + P * result after every call, with x10 / x100 gradient-scale jumps, in the
    lossless layout and with slot_factor=1 (overflow policy): nothing selected is ever lost, replicas stay identical."""
    if torch.cuda.device_count() < P:
        pytest.skip("needs %d GPUs" % P)
    n, iters = 500_003, 10
    kw = dict(density=0.01, local_recompute_interval=8, global_recompute_interval=8, repartition_interval=8,
              slot_factor=slot_factor, gather_factor=0.0)
    got = run_distributed(_conservation_worker, P, (n, iters, kw), backend="nccl", timeout=600)
    redo = 0
    for it in range(iters):
        tot_acc = sum(got[r][0][it] for r in range(P))
        tot_res = sum(got[r][2][it] for r in range(P))
        out = got[0][1][it].double()
        for r in range(1, P):
            assert torch.equal(got[r][1][it], got[0][1][it]), "replicas differ at it %d" % it
        err = (tot_acc - tot_res - P * out).abs().max().item()
        scale = tot_acc.abs().max().item()
        assert err <= 1e-5 * scale, "it %d: mass not conserved: err %g (scale %g) stats %s" % (it, err, scale, got[0][3][it])
        for r in range(P):
            st = got[r][3][it]
            assert st["overflow_send"] == 0 and st["overflow_gather"] == 0 and st["fault"] == 0, (it, r, st)
            redo += st["redo"]
    if slot_factor > 0:
        assert redo > 0, "the overflow policy never ran"


def _stress_worker(rank, P, n, calls):
    """Flag-protocol stress: thousands of back-to-back calls, random per-rank delays between them (both mailbox
    parities, ranks arriving in every order), replicas compared at the end."""
    import random
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine
    from oktopk_b200.parallel.world import World
    w = World()
    cfg = OkTopkConfig(density=0.01, local_recompute_interval=16, global_recompute_interval=16, repartition_interval=32,
                       peer_timeout_s=30.0)
    eng = CudaBucketEngine(n, cfg, w, name="t")
    rng = random.Random(1234 + rank)
    base = torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(rank))
    chk = 0.0
    for it in range(calls):
        eng.grad.copy_(base)
        eng.grad.mul_(1.0 + 0.001 * (it % 7))
        if rng.random() < 0.3:
            torch.cuda._sleep(int(rng.random() * 200_000))           # up to ~100 us of extra delay on this rank only
        eng.reduce("oktopk")
        if it % 1000 == 999:
            torch.cuda.synchronize()
            chk += float(eng.grad.double().sum())
    torch.cuda.synchronize()
    st = eng.stats()
    final = eng.grad.cpu().clone()
    w.barrier()
    eng.close()
    return chk, final, st["fault"], st["cum_overflow_send"], st["epoch"]


@pytest.mark.parametrize("P", [2, 8])
def test_flag_protocol_stress(P):
    if torch.cuda.device_count() < P:
        pytest.skip("needs %d GPUs" % P)
    calls = 10_000
    got = run_distributed(_stress_worker, P, (65_536, calls), backend="nccl", timeout=900)
    for r in range(P):
        assert got[r][2] == 0 and got[r][3] == 0 and got[r][4] == calls, got[r][2:]
        assert got[r][0] == got[0][0]
        assert torch.equal(got[r][1], got[0][1])


def _nvls_worker(rank, P, n):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine
    from oktopk_b200.parallel.symm import _NVLS_STATE
    from oktopk_b200.parallel.world import World
    w = World()
    eng = CudaBucketEngine(n, OkTopkConfig(compressor="none", sparse=False), w, name="t")
    outs = []
    for it in range(3):
        eng.grad.copy_(_grad(it, rank, n).cuda())
        torch.cuda.synchronize()
        w.barrier()
        eng.reduce("none")
        torch.cuda.synchronize()
        outs.append(eng.grad.cpu().clone())
    st = eng.stats()
    w.barrier()
    eng.close()
    return outs, st["nvls"], dict(_NVLS_STATE), st["fault"]


@pytest.mark.parametrize("P", [2, 8])
def test_dense_allreduce_nvls_or_peer_path(P):
    """The dense kernel (multimem path when the box exposes an NVSwitch multicast object, peer loads otherwise)
    against a plain fp32 sum."""
    if torch.cuda.device_count() < P:
        pytest.skip("needs %d GPUs" % P)
    n = 1_000_003
    got = run_distributed(_nvls_worker, P, (n,), backend="nccl", timeout=240)
    print("NVLS:", got[0][1], got[0][2])
    for it in range(3):
        ref = sum(_grad(it, r, n).double() for r in range(P)) / P
        for r in range(P):
            assert got[r][3] == 0
            torch.testing.assert_close(got[r][0][it].double(), ref, rtol=1e-5, atol=1e-6)
            assert torch.equal(got[r][0][it], got[0][0][it])


def test_automatic_dense_switch_at_high_density():
    """density >= dense_switch_density: the engine reduces the error-compensated gradient with the dense kernel
    (sum of (g + residual) / P everywhere, residual cleared)."""
    P, n = 2, 200_000
    kw = dict(density=0.1, dense_switch_density=0.05)
    got = run_distributed(_engine_worker, P, ("oktopk", n, 2, kw), backend="nccl", timeout=300)
    acc = [torch.zeros(n) for _ in range(P)]
    for it in range(2):
        ref = sum((_grad(it, r, n) + acc[r]).double() for r in range(P)) / P
        for r in range(P):
            torch.testing.assert_close(got[r][0][it].double(), ref, rtol=1e-5, atol=1e-6)
            assert got[r][2][it]["mode"] == "dense(auto)"
        acc = [torch.zeros(n) for _ in range(P)]
    for r in range(P):
        assert float(got[r][1].abs().max()) == 0.0


def test_small_bucket_for_sanitizer_runs():
    """A deliberately tiny 2-GPU Ok-Topk run (all flavours, both slot layouts): the target of scripts/sanitize.sh's
    synccheck / racecheck / memcheck passes over the cross-GPU path."""
    for sf in (0.0, 2.0):
        _check("oktopk", 2, 65_536, 5, dict(density=0.01, local_recompute_interval=2, global_recompute_interval=2,
                                           repartition_interval=2, slot_factor=sf, gather_factor=sf))
    _check("gtopk", 2, 65_536, 2, dict(density=0.01))
