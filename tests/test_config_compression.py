"""Config presets and compressor semantics -- CPU only."""
import math

import pytest
import torch

import oktopk_b200 as okt
from oktopk_b200 import compression as C
from oktopk_b200.config import OkTopkConfig, preset, sigma_scale_for_density


def test_presets_carry_the_reference_literals():
    v, l, b = preset("vgg16"), preset("lstm_an4"), preset("bert_base")
    assert (v.warmup_iters, l.warmup_iters, b.warmup_iters) == (512, 128, 0)
    assert (v.local_recompute_interval, l.local_recompute_interval, b.local_recompute_interval) == (32, 32, 128)
    assert (v.global_recompute_interval, b.global_recompute_interval) == (32, 128)
    assert v.repartition_interval == l.repartition_interval == b.repartition_interval == 64
    assert (v.overselect_guard_num, v.overselect_guard_den, v.overselect_guard_loops) == (4, 3, 5)
    assert (l.overselect_guard_num, l.overselect_guard_den) == (3, 2)
    assert b.overselect_guard_loops == 0
    assert v.local_adapt_factor == l.local_adapt_factor == 1.012 and b.local_adapt_factor == 1.025
    assert (v.global_adapt_inc, v.global_adapt_dec) == (1.008, 1.008)
    assert (l.global_adapt_inc, l.global_adapt_dec) == (1.01, 1.008)
    assert (b.global_adapt_inc, b.global_adapt_dec) == (1.036, 1.025)
    assert b.balanced_allgather and not v.balanced_allgather
    assert v.density == l.density == 0.02 and b.density == 0.01
    assert preset("bert", density=0.001).density == 0.001
    with pytest.raises(KeyError):
        preset("nope")


def test_sigma_scale_table():
    assert sigma_scale_for_density(0.8) == 0.5
    assert sigma_scale_for_density(0.1) == 1.5
    assert sigma_scale_for_density(0.02) == 2.0
    assert sigma_scale_for_density(0.001) == 3.0


def test_registry_has_every_reference_name():
    want = {"topkA", "topkAopt", "topkA2", "topkSA", "gtopk", "gaussiank", "gaussiankconcat", "gaussiankSA", "oktopk",
            "none"}
    assert want <= set(k for k in C.compressors if k)
    for name in want:
        inst = C.resolve_compressor(name)
        assert inst.name == name
    # oktopk / topkAopt inherit the Gaussian family, topkA/gtopk the TopK family (VGG/compression.py:484-509)
    assert issubclass(C.compressors["oktopk"], C.GaussianCompressor)
    assert issubclass(C.compressors["topkAopt"], C.GaussianCompressor)
    assert issubclass(C.compressors["gtopk"], C.TopKCompressor)
    assert issubclass(C.compressors["topkA"], C.TopKCompressor)
    assert okt.compressors is C.compressors


def test_none_compressor_is_identity():
    t = torch.randn(10)
    out, ctx = C.NoneCompressor.compress(t)
    assert out is t and ctx is None
    assert C.NoneCompressor.decompress(t) is t


def test_topk_compress_org_error_feedback():
    torch.manual_seed(0)
    c = C.TopKCompressor()
    n, ratio = 1000, 0.05
    g1 = torch.randn(n)
    t = g1.clone()
    out, idx = c.compress_org(t, "b", ratio)
    k = int(n * ratio)
    assert idx.numel() == k
    res = c.residual("b", t)
    # selected entries keep their value in the tensor and are zero in the residual; the rest is the residual
    assert int((out != 0).sum()) == k
    torch.testing.assert_close(out + res, g1)
    assert float(res[idx].abs().max()) == 0.0
    # the k kept are the k largest magnitudes
    assert float(out[idx].abs().min()) >= float(res.abs().max())
    # second call accumulates the residual first
    g2 = torch.randn(n)
    t2 = g2.clone()
    out2, idx2 = c.compress_org(t2, "b", ratio)
    torch.testing.assert_close(out2 + c.residual("b", t2), g1 - out + g2)


def test_topk_ratio2threshold_zeroes_topk_in_residual():
    torch.manual_seed(1)
    c = C.TopKCompressor()
    g = torch.randn(2000)
    t = g.clone()
    thr = c.ratio2threshold(t, "x", 0.01)
    k = 20
    ref = float(torch.topk(g.abs(), k).values[-1])
    assert thr == pytest.approx(ref)
    res = c.residual("x", t)
    assert int((res == 0).sum()) >= k
    assert float(res.abs().max()) <= thr


def test_add_residuals_puts_back_the_losers():
    torch.manual_seed(2)
    c = C.TopKCompressor()
    g = torch.randn(500)
    t = g.clone()
    out, idx = c.compress_org(t, "z", 0.1)
    included = torch.arange(0, idx.numel(), 2)              # every other local pick survived globally
    c.add_residuals(included, "z")
    res = c.residual("z", t)
    losers = idx[torch.arange(1, idx.numel(), 2)]
    torch.testing.assert_close(res[losers], g[losers])
    assert float(res[idx[included]].abs().max()) == 0.0


def test_gaussian_ratio2threshold_keeps_full_accumulator():
    torch.manual_seed(3)
    c = C.GaussianCompressor()
    g = torch.randn(4000)
    t = g.clone()
    thr = c.ratio2threshold(t, "g", 0.01)
    assert thr == pytest.approx(float(torch.topk(g.abs(), 40).values[-1]))
    torch.testing.assert_close(c.residual("g", t), g)       # nothing zeroed (VGG/compression.py:370-381)
    # add2residual accumulates and applies the over-selection guard (thr*1.03 up to 5x while cnt > 4k/3)
    g2 = torch.randn(4000) * 3
    t2 = g2.clone()
    thr2 = c.add2residual(t2, "g", thr, 40)
    acc = g + g2
    torch.testing.assert_close(c.residual("g", t2), acc)
    expect = thr
    for _ in range(5):
        if int((acc.abs() > expect).sum()) > 4 * 40 / 3:
            expect *= 1.03
    assert thr2 == pytest.approx(expect, rel=1e-6)


def test_compressbythreshold_is_strict_and_int32():
    t = torch.tensor([0.5, -1.0, 1.0, 2.0, -3.0])
    idx, val = C.Compressor.compressbythreshold(t, 1.0)
    assert idx.dtype == torch.int32
    assert idx.tolist() == [3, 4] and val.tolist() == [2.0, -3.0]
    idl = C.Compressor.compressbythresholdlong(t, 0.4)
    idl = idl[0] if isinstance(idl, tuple) else idl
    assert idl.dtype == torch.int64 and idl.numel() == 5


def test_k2globalthreshold():
    v = torch.tensor([0.1, -5.0, 3.0, -0.2, 4.0])
    vals, pos, thr = C.Compressor.k2globalthreshold(v, 2)
    assert sorted(pos.tolist()) == [1, 4] and thr == pytest.approx(4.0)
    vals, pos, thr = C.Compressor.k2globalthreshold(v, 50)       # kk = min(len, k)
    assert pos.numel() == 5 and thr == pytest.approx(0.1)


def test_gaussian_threshold_formula():
    lo, hi = C.gen_threshold_from_normal_distribution(1 - 0.01, 0.0, 1.0)
    # right tail mass rho/2 on each side
    from scipy import stats
    assert hi == pytest.approx(stats.norm.ppf(1 - 0.005), rel=1e-6)
    assert lo == pytest.approx(-hi, rel=1e-6)


def test_gaussian_compress_selects_about_k():
    torch.manual_seed(4)
    cfg = OkTopkConfig(density=0.01)
    c = C.GaussianCompressor(cfg)
    g = torch.randn(100_000)
    t = g.clone()
    idx, val = c.compress(t, "q", 0.01)
    k = 1000
    assert 0.75 * k <= idx.numel() <= 1.25 * k
    res = c.residual("q", t)
    assert float(res[idx.long()].abs().max()) == 0.0
    torch.testing.assert_close(val, g[idx.long()])


def test_compressor_state_is_per_instance_not_class_level():
    a, b = C.TopKCompressor(), C.TopKCompressor()
    t = torch.ones(8)
    a.compress_org(t.clone(), "n", 0.5)
    assert float(b.residual("n", t).abs().sum()) == 0.0        # the reference shares class-level dicts (A.4-3)
    a.clear()
    assert float(a.residual("n", t).abs().sum()) == 0.0


def test_config_roundtrip_and_k():
    cfg = OkTopkConfig(density=0.001)
    assert cfg.k_for(14_728_266) == 14728
    d = cfg.to_dict()
    assert OkTopkConfig(**d) == cfg
    assert math.isclose(cfg.replace(density=0.5).density, 0.5)
