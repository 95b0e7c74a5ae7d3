"""Property tests of the single-process oracle (what the new framework must bring itself).

The oracle simulates P ranks on a list of tensors and encodes SURVEY 3.3 / Appendix B exactly; the CUDA
kernels and the torch.distributed path are both checked against it elsewhere.  Here: invariants.
"""
import pytest
import torch

from oktopk_b200.config import OkTopkConfig
from oktopk_b200.parallel.oracle import (ORACLES, adapt_global, adapt_local, boundaries_from_cuts, kth_largest_abs,
                                         merge_topk, quantile_cuts, run_oracle)
from oktopk_b200.parallel.state import SparseState, offsets_of, uniform_boundaries


def _grads(P, n, it, scale=1.0):
    out = []
    for r in range(P):
        g = torch.Generator().manual_seed(100 * it + r)
        out.append(torch.randn(n, generator=g) * scale * torch.linspace(0.3, 1.7, n))
    return out


def test_kth_largest_abs_and_cuts():
    x = torch.tensor([1.0, -7.0, 3.0, -2.0, 5.0])
    assert kth_largest_abs(x, 1) == 7.0 and kth_largest_abs(x, 3) == 3.0
    sel = torch.tensor([2, 5, 9, 11, 30, 31, 50, 77])
    assert quantile_cuts(sel, 4, 100) == [9, 30, 50]          # I[j*|I|//P]
    b, off = boundaries_from_cuts([9, 30, 50], 100)
    assert b == [9, 21, 20, 50] and off == [0, 9, 30, 50] and sum(b) == 100
    # non-monotone averaged cuts are clamped instead of tripping the reference's assert (A.4-5)
    b, off = boundaries_from_cuts([40, 30, 90], 100)
    assert sum(b) == 100 and all(x >= 0 for x in b)
    assert uniform_boundaries(10, 4) == [2, 2, 2, 4] and offsets_of([2, 2, 2, 4]) == [0, 2, 4, 6]


def test_threshold_adaptation_rules():
    cfg = OkTopkConfig()        # VGG numbers: 2k/3, 5k/4, x1.012 ; 2k/3 -> /1.008, 4k/3 -> x1.008
    k = 300
    assert adapt_local(1.0, 100, k, cfg) == pytest.approx(1 / 1.012)
    assert adapt_local(1.0, 300, k, cfg) == 1.0
    assert adapt_local(1.0, 400, k, cfg) == pytest.approx(1.012)
    assert adapt_global(1.0, 100, k, cfg) == pytest.approx(1 / 1.008)
    assert adapt_global(1.0, 401, k, cfg) == pytest.approx(1.008)


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_oktopk_invariants(P):
    n, iters = 20_000, 12
    cfg = OkTopkConfig(density=0.01, local_recompute_interval=4, global_recompute_interval=4, repartition_interval=8)
    k = int(n * cfg.density)
    states = [SparseState(n, P) for _ in range(P)]
    for it in range(iters):
        grads = _grads(P, n, it)
        acc_before = [g + (st.residual if st.residual is not None else 0) for g, st in zip(grads, states)]
        thr_before = [st.local_thr for st in states]
        out = run_oracle("oktopk", [g.clone() for g in grads], states, cfg)
        res = out[0]
        # every rank ends with the same dense result
        for r in range(1, P):
            assert torch.equal(out[r], res)
        # support of the result is inside the union of the local selections
        union = torch.zeros(n, dtype=torch.bool)
        for r in range(P):
            # locally selected <=> residual was cleared there or |acc| above the threshold used
            union |= acc_before[r].abs() > 0
        nz = res != 0
        assert int(nz.sum()) == states[0].last_global_count
        assert bool((nz & ~union).sum() == 0)
        # conservation of gradient mass per rank: acc_before == residual_after + (what this rank contributed & was kept)
        for r in range(P):
            contributed = acc_before[r] - states[r].residual
            # contributions exist only at globally kept positions
            assert bool(((contributed != 0) & ~nz).sum() == 0)
        # the result is exactly the mean of the kept contributions
        total = sum(acc_before[r] - states[r].residual for r in range(P)) / P
        # kept positions may also hold contributions whose residual was NOT cleared (not locally above thr): none by
        # construction, so the sums agree on the support
        torch.testing.assert_close(res[nz], total[nz], rtol=1e-5, atol=1e-6)
        # exact iterations keep at most k (+ties) entries globally
        if it % cfg.global_recompute_interval == 0:
            assert states[0].last_global_count <= k + 2
        # communication volume: < 6k(P-1)/P scalars... when selections are ~k (generous factor for stale thresholds)
        # (exact-global iterations gather every reduced non-zero, up to P*k: the bound is for threshold-reuse steps)
        if P > 1 and it % cfg.global_recompute_interval != 0:
            for st in states:
                assert st.last_volume_elems <= 6 * 2.0 * max(st.last_local_count, states[0].last_global_count, k)
        # regions tile the bucket
        assert sum(states[0].boundaries) == n and states[0].region_offsets[0] == 0


def test_oktopk_p1_exact_iteration_is_plain_topk():
    n = 5000
    cfg = OkTopkConfig(density=0.02)
    st = [SparseState(n, 1)]
    g = _grads(1, n, 0)[0]
    out = run_oracle("oktopk", [g.clone()], st, cfg)[0]
    k = int(n * 0.02)
    thr = float(torch.topk(g.abs(), k).values[-1])
    # strict '>' select with the exact k-th value as threshold: k-1 entries (+ties excluded), SURVEY 3.3 note
    keep = g.abs() > thr
    assert torch.equal(out != 0, keep)
    torch.testing.assert_close(out[keep], g[keep])
    torch.testing.assert_close(st[0].residual, torch.where(keep, torch.zeros_like(g), g))


@pytest.mark.parametrize("name", ["topkA", "topkA2", "topkAopt", "gaussiank", "gaussiankconcat", "topkSA", "gaussiankSA",
                                  "gtopk", "none"])
@pytest.mark.parametrize("P", [2, 4])
def test_every_scheme_runs_and_agrees_across_ranks(name, P):
    n = 8192
    cfg = OkTopkConfig(density=0.02, topkaopt_recompute_interval=3)
    states = [SparseState(n, P) for _ in range(P)]
    for it in range(5):
        grads = _grads(P, n, it)
        dense = torch.stack(grads).sum(0) / P
        out = run_oracle(name, [g.clone() for g in grads], states, cfg)
        for r in range(1, P):
            assert torch.equal(out[r], out[0]), name
        if name == "none":
            torch.testing.assert_close(out[0], dense)
            continue
        nnz = int((out[0] != 0).sum())
        assert 0 < nnz < n // 2, (name, nnz)           # sparse (stale-threshold schemes over-select, as the reference)
        assert all(st.counter == it + 1 for st in states)
        assert states[0].last_mode != ""


def test_topka_result_is_mean_of_local_topk():
    P, n = 4, 4096
    cfg = OkTopkConfig(density=0.01)
    states = [SparseState(n, P) for _ in range(P)]
    grads = _grads(P, n, 0)
    k = int(n * cfg.density)
    ref = torch.zeros(n)
    for g in grads:
        idx = torch.topk(g.abs(), k).indices
        ref[idx] += g[idx]
    ref /= P
    out = run_oracle("topkA", [g.clone() for g in grads], states, cfg)[0]
    torch.testing.assert_close(out, ref)
    for g, st in zip(grads, states):
        idx = torch.topk(g.abs(), k).indices
        assert float(st.residual[idx].abs().max()) == 0.0


def test_gtopk_keeps_k_and_restores_losers():
    P, n = 4, 4096
    cfg = OkTopkConfig(density=0.01)
    k = int(n * cfg.density)
    states = [SparseState(n, P) for _ in range(P)]
    grads = _grads(P, n, 1)
    out = run_oracle("gtopk", [g.clone() for g in grads], states, cfg)[0]
    kept = out != 0
    assert int(kept.sum()) <= k
    for g, st in zip(grads, states):
        idx = torch.topk(g.abs(), k).indices
        mine = torch.zeros(n, dtype=torch.bool)
        mine[idx] = True
        # local picks that did not survive the tree go back into the residual (VGG/compression.py:151-160)
        lost = mine & ~kept
        torch.testing.assert_close(st.residual[lost], g[lost])
        assert float(st.residual[mine & kept].abs().max() if (mine & kept).any() else 0.0) == 0.0


def test_merge_topk_adds_coincident_indices():
    a = (torch.tensor([1, 5, 9]), torch.tensor([1.0, -4.0, 2.0]))
    b = (torch.tensor([5, 7]), torch.tensor([-1.0, 3.0]))
    idx, val = merge_topk(a, b, 2, 16)
    d = dict(zip(idx.tolist(), val.tolist()))
    assert d == {5: -5.0, 7: 3.0}


def test_warmup_is_dense_and_leaves_sparse_state_untouched():
    P, n = 2, 1000
    cfg = OkTopkConfig(density=0.02, warmup_iters=3)
    states = [SparseState(n, P) for _ in range(P)]
    for it in range(3):
        grads = _grads(P, n, it)
        out = run_oracle("oktopk", [g.clone() for g in grads], states, cfg)
        torch.testing.assert_close(out[0], (grads[0] + grads[1]) / 2)
        assert states[0].last_mode == "dense" and states[0].residual is None
    run_oracle("oktopk", _grads(P, n, 3), states, cfg)
    assert states[0].last_mode == "oktopk" and states[0].local_thr > 0          # first sparse step is an exact one (A.4-4)


def test_state_dict_roundtrip_resumes_identically():
    P, n = 2, 6000
    cfg = OkTopkConfig(density=0.01, local_recompute_interval=4, global_recompute_interval=4, repartition_interval=4)
    A = [SparseState(n, P) for _ in range(P)]
    for it in range(5):
        run_oracle("oktopk", _grads(P, n, it), A, cfg)
    B = [SparseState(n, P) for _ in range(P)]
    for a, b in zip(A, B):
        b.load_state_dict(a.state_dict())
    for it in range(5, 9):
        oa = run_oracle("oktopk", _grads(P, n, it), A, cfg)
        ob = run_oracle("oktopk", _grads(P, n, it), B, cfg)
        assert torch.equal(oa[0], ob[0])
    assert set(ORACLES) >= {"oktopk", "topkA", "topkA2", "topkAopt", "topkSA", "gtopk", "gaussiank", "gaussiankSA"}


@pytest.mark.parametrize("n", [3, 10, 65])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_tiny_buckets_are_handled_by_every_scheme(n, P):
    """Buckets smaller than the world / k = 1 (bias-only buckets happen with small ``bucket_elems``)."""
    for name in ["oktopk", "topkA", "topkA2", "topkAopt", "topkSA", "gaussiank", "gaussiankSA", "gtopk", "none"]:
        cfg = OkTopkConfig(density=0.001, local_recompute_interval=2, global_recompute_interval=2, repartition_interval=2)
        st = [SparseState(n, P) for _ in range(P)]
        for it in range(4):
            g = [torch.randn(n, generator=torch.Generator().manual_seed(it * 10 + r)) for r in range(P)]
            out = run_oracle(name, g, st, cfg)
            assert all(torch.isfinite(o).all() for o in out), (name, n, P)
            assert all(torch.equal(o, out[0]) for o in out)


@pytest.mark.parametrize("P", [1, 4])
def test_overselect_cap_bounds_the_volume_and_conserves_mass(P):
    """overselect_cap = 2: after x10 / x100 gradient-scale jumps the stale threshold climbs the coarse ladder rungs until
    at most 2k entries per rank are selected; nothing is lost (sum acc == sum residual + P * result)."""
    n, iters = 20_000, 12
    cfg = OkTopkConfig(density=0.01, local_recompute_interval=8, global_recompute_interval=8, repartition_interval=8,
                       overselect_cap=2.0)
    ref = cfg.replace(overselect_cap=0.0)
    k = int(n * cfg.density)
    states = [SparseState(n, P) for _ in range(P)]
    states0 = [SparseState(n, P) for _ in range(P)]
    worst_uncapped = 0
    for it in range(iters):
        scale = 1.0 if it < 3 else (10.0 if it < 6 else 100.0)
        grads = _grads(P, n, it, scale)
        acc = sum((g + (st.residual if st.residual is not None else 0)).double() for g, st in zip(grads, states))
        out = run_oracle("oktopk", [g.clone() for g in grads], states, cfg)
        run_oracle("oktopk", [g.clone() for g in grads], states0, ref)
        worst_uncapped = max(worst_uncapped, max(st.last_local_count for st in states0))
        if it % 8 != 0:                                   # threshold-reuse iterations
            assert all(st.last_local_count <= 2 * k for st in states), (it, [st.last_local_count for st in states])
        tot_res = sum(st.residual.double() for st in states)
        err = (acc - tot_res - P * out[0].double()).abs().max().item()
        assert err <= 1e-5 * max(acc.abs().max().item(), 1.0), (it, err)
    assert worst_uncapped > 10 * k                        # without the cap the same stream over-selects massively
