"""Multi-process CPU tests (gloo, 127.0.0.1): the torch.distributed implementation of every scheme must
reproduce the single-process oracle, and the optimizer wrapper must keep replicas identical.

This is BASELINE.json's config #1 (plumbing without a GPU) and SURVEY 4's "multi-node without a cluster".
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mp_util import run_distributed  # noqa: E402


def _grad(it, rank, n):
    g = torch.Generator().manual_seed(1000 * it + rank)
    return torch.randn(n, generator=g) * torch.linspace(0.2, 2.0, n)


def _algo_worker(rank, P, name, n, iters, cfg_kw):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.algorithms import sparse_allreduce
    from oktopk_b200.parallel.state import SparseState
    from oktopk_b200.parallel.world import World
    w = World()
    cfg = OkTopkConfig(**cfg_kw)
    st = SparseState(n, P)
    outs, meta = [], []
    for it in range(iters):
        g = _grad(it, rank, n)
        sparse_allreduce(name, g, st, cfg, w)
        outs.append(g.clone())
        meta.append((st.local_thr, st.global_thr, list(st.region_offsets), st.last_local_count, st.last_global_count,
                     st.last_volume_elems))
    return outs, (st.residual.clone() if st.residual is not None else None), meta


def _check(name, P, n=6000, iters=9, exact=True, **cfg_kw):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.oracle import run_oracle
    from oktopk_b200.parallel.state import SparseState
    cfg_kw = dict(dict(density=0.02, local_recompute_interval=4, global_recompute_interval=4, repartition_interval=4,
                       topkaopt_recompute_interval=3), **cfg_kw)
    got = run_distributed(_algo_worker, P, (name, n, iters, cfg_kw), backend="gloo", timeout=240)
    cfg = OkTopkConfig(**cfg_kw)
    states = [SparseState(n, P) for _ in range(P)]
    for it in range(iters):
        ref = run_oracle(name, [_grad(it, r, n) for r in range(P)], states, cfg)
        for r in range(P):
            if exact:
                assert torch.equal(got[r][0][it], ref[r]), "%s it %d rank %d differs from the oracle" % (name, it, r)
            else:
                torch.testing.assert_close(got[r][0][it], ref[r], rtol=1e-5, atol=1e-6)
        if name == "oktopk":
            for r in range(P):
                lt, gt, off, lc, gc, vol = got[r][2][it]
                assert off == states[r].region_offsets
                assert lc == states[r].last_local_count and gc == states[r].last_global_count
                assert lt == pytest.approx(states[r].local_thr, rel=1e-6)
                assert gt == pytest.approx(states[r].global_thr, rel=1e-6)
    for r in range(P):
        if states[r].residual is not None:
            assert torch.equal(got[r][1], states[r].residual), "residual of rank %d differs" % r
    return got


def test_oktopk_gloo_world2_matches_oracle():
    got = _check("oktopk", 2)
    # volume accounting: threshold-reuse steps stay within the 6k(P-1)/P bound (README.md:2 of the reference)
    k = int(6000 * 0.02)
    for it, m in enumerate(got[0][2]):
        if it % 4 != 0:
            assert m[5] <= 6 * max(m[3], m[4], k)


def test_oktopk_gloo_world4_matches_oracle():
    _check("oktopk", 4, n=5000, iters=6, exact=False)       # 4-way float sums: order may differ from the oracle's


@pytest.mark.parametrize("preset", ["lstm_an4", "bert_base"])
def test_oktopk_presets_gloo_world2(preset):
    import dataclasses
    import oktopk_b200 as okt
    kw = dataclasses.asdict(okt.preset(preset, density=0.02, warmup_iters=1, local_recompute_interval=3,
                                       global_recompute_interval=3, repartition_interval=4))
    _check("oktopk", 2, iters=7, **kw)


@pytest.mark.parametrize("name", ["topkA", "topkA2", "topkAopt", "topkSA", "gaussiankSA", "gtopk", "gaussiank",
                                  "gaussiankconcat", "none"])
def test_baseline_schemes_gloo_world2_match_oracle(name):
    _check(name, 2, iters=5)


def test_gtopk_gloo_world4_tree():
    _check("gtopk", 4, n=4096, iters=3, exact=False)


def test_topkdsa_dense_fallback_gloo():
    # density high enough that the reduced regions hold >= n/3 non-zeros => dense Allgatherv branch (VGG:1346-1353)
    _check("topkSA", 2, n=3000, iters=3, density=0.4, dense_switch_density=0.0)


def test_dense_switch_gloo_matches_oracle():
    """density >= dense_switch_density: the dist path and the oracle both reduce (gradient + residual) densely."""
    got = _check("oktopk", 2, n=3000, iters=4, density=0.1, dense_switch_density=0.05, exact=False)
    assert got[0][2][-1] is not None


# --------------------------------------------------------------------------------------------- optimizer wrapper
def _opt_worker(rank, P, compressor, density, steps):
    import oktopk_b200 as okt
    torch.manual_seed(0)                                     # identical init on every rank
    net = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5))
    ref = None
    if compressor == "none":
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5))
        ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    cfg = okt.OkTopkConfig(density=density, local_recompute_interval=2, global_recompute_interval=2,
                           repartition_interval=2, bucket_elems=300)
    opt = okt.DistributedOptimizer(torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4),
                                   named_parameters=net.named_parameters(), compression=okt.compressors[compressor],
                                   is_sparse=compressor != "none", cfg=cfg)
    assert okt.optimizer.rank() == rank and okt.optimizer.size() == P
    losses = []
    for it in range(steps):
        g = torch.Generator().manual_seed(10 * it + rank)
        x, y = torch.randn(8, 20, generator=g), torch.randint(0, 5, (8,), generator=g)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        if ref is not None:                                  # dense: must equal SGD on the rank-averaged gradient
            ref_opt.zero_grad()
            tot = 0
            for r in range(P):
                gg = torch.Generator().manual_seed(10 * it + r)
                xr, yr = torch.randn(8, 20, generator=gg), torch.randint(0, 5, (8,), generator=gg)
                tot = tot + torch.nn.functional.cross_entropy(ref(xr), yr) / P
            tot.backward()
            ref_opt.step()
    flat = torch.cat([p.detach().view(-1) for p in net.parameters()])
    refflat = torch.cat([p.detach().view(-1) for p in ref.parameters()]) if ref is not None else None
    stats = opt.comm_stats()
    nb = len(stats)
    opt.close()
    return flat, refflat, losses, nb


def test_distributed_optimizer_dense_gloo_equals_sgd_on_mean_gradient():
    out = run_distributed(_opt_worker, 2, ("none", 1.0, 5), backend="gloo")
    assert torch.equal(out[0][0], out[1][0])
    torch.testing.assert_close(out[0][0], out[0][1], rtol=1e-5, atol=1e-6)
    assert out[0][3] >= 2                                    # bucket_elems=300 => several buckets


def test_distributed_optimizer_oktopk_gloo_replicas_identical():
    out = run_distributed(_opt_worker, 2, ("oktopk", 0.05, 8), backend="gloo")
    assert torch.equal(out[0][0], out[1][0])
    assert all(torch.isfinite(out[r][0]).all() for r in range(2))


# --------------------------------------------------------------------------------------------- BASELINE config #1
def _vgg_worker(rank, P, steps):
    from oktopk_b200.train.trainer import Trainer
    import oktopk_b200 as okt
    cfg = okt.preset("vgg16", density=0.01, warmup_iters=1, local_recompute_interval=2, global_recompute_interval=2,
                     repartition_interval=2)
    tr = Trainer(dnn="vgg16", dataset="cifar10", batch_size=2, lr=0.01, compressor="oktopk", density=0.01, cfg=cfg,
                 device=torch.device("cpu"))
    for _ in range(steps):
        tr.train_step()
    loss = tr.last_loss()
    st = tr.optimizer.comm_stats()
    flat = torch.cat([p.detach().view(-1) for p in tr.net.parameters()])
    dens = tr.optimizer.get_current_density()
    tr.close()
    return flat[::97].clone(), loss, {k: (v["mode"], v["local_count"], v["global_count"]) for k, v in st.items()}, dens


def test_vgg16_oktopk_density001_gloo_world2_plumbing():
    """BASELINE.json configs[0]: VGG-16 Ok-Topk density=0.01 on CPU/gloo world_size=2."""
    out = run_distributed(_vgg_worker, 2, (3,), backend="gloo", timeout=600)
    assert torch.equal(out[0][0], out[1][0]), "replicas diverged"
    assert all(torch.isfinite(torch.tensor(o[1])) for o in out)
    modes = list(out[0][2].values())
    assert modes and all(m[0] == "oktopk" for m in modes), modes        # past the 1-step dense warm-up
    k = int(14_728_266 * 0.01)
    assert 0 < modes[0][2] <= 3 * k
    assert out[0][3] == pytest.approx(0.01)


def _tiny_worker(rank, P, n):
    from oktopk_b200.config import OkTopkConfig
    from oktopk_b200.parallel.algorithms import sparse_allreduce
    from oktopk_b200.parallel.state import SparseState
    from oktopk_b200.parallel.world import World
    w = World()
    outs = {}
    for name in ["oktopk", "topkA", "topkA2", "topkAopt", "topkSA", "gaussiank", "gaussiankSA", "gtopk", "none"]:
        cfg = OkTopkConfig(density=0.001, local_recompute_interval=2, global_recompute_interval=2, repartition_interval=2)
        st = SparseState(n, P)
        for it in range(3):
            g = torch.randn(n, generator=torch.Generator().manual_seed(it * 10 + rank))
            sparse_allreduce(name, g, st, cfg, w)
        outs[name] = g.clone()
    return outs


def test_tiny_bucket_gloo_world2_all_schemes():
    out = run_distributed(_tiny_worker, 2, (10,), backend="gloo", timeout=120)
    for name, t in out[0].items():
        assert torch.isfinite(t).all() and torch.equal(t, out[1][name]), name


# --------------------------------------------------------------------------------------------- per-rank checkpoints
def _ckpt_worker(rank, P, directory):
    import os
    import oktopk_b200 as okt
    from oktopk_b200.train.trainer import Trainer
    cfg = okt.preset("vgg16", density=0.02, warmup_iters=1, local_recompute_interval=2, global_recompute_interval=2)
    mk = lambda: Trainer(dnn="mnistnet", dataset="mnist", batch_size=4, lr=0.05, compressor="oktopk", density=0.02, cfg=cfg,
                         device=torch.device("cpu"))
    tr = mk()
    for _ in range(4):
        tr.train_step()
    path = os.path.join(directory, "mnistnet-rank0-epoch0.pth")
    tr.save_checkpoint(path)                                   # collective: every rank writes its own sparse state
    res = [st.residual.clone() for st in tr.optimizer._allreducer._dist_states.values()]
    tr.close()
    tr2 = mk()
    tr2.load_checkpoint(path)
    res2 = [st.residual.clone() for st in tr2.optimizer._allreducer._dist_states.values()]
    tr2.close()
    return res, res2, os.path.isfile("%s.rank%d" % (path, rank))


def test_checkpoint_keeps_every_ranks_own_residual(tmp_path):
    """Error-feedback residuals are per-rank: a resume must give each rank ITS residual back (not rank 0's)."""
    out = run_distributed(_ckpt_worker, 2, (str(tmp_path),), backend="gloo", timeout=300)
    for r in range(2):
        assert out[r][2]
        assert len(out[r][0]) == len(out[r][1]) >= 1
        for a, b in zip(out[r][0], out[r][1]):
            assert torch.equal(a, b)
    assert any(not torch.equal(a, b) for a, b in zip(out[0][0], out[1][0])), "the two ranks' residuals should differ"
