#!/usr/bin/env python
"""Headline benchmark: training throughput (samples/s, whole job) of a reference workload with the
Ok-Topk sparse allreduce at density 0.001 -- BASELINE.json's metric/config.

    python bench.py [--gpus N --steps K --warmup W] [--model vgg16|lstman4|bert] [--impl ours|reference]

N > 1 is launched by the driver through ``torch.distributed.run`` (one rank per GPU, NCCL bootstrap;
the sparse allreduce itself runs on the fused peer-memory kernels).  Synthetic data of the named shape,
random-init weights, fp32 compute (the reference's precision).  Prints ONE JSON line on rank 0.

Two numbers:
  * ``value``  - device-timed (CUDA events, max over ranks) over exactly K optimizer steps with the
                 batch already resident on the device;
  * ``e2e``    - the same metric through the public API (``Trainer.train_step``): every step copies its
                 batch host->device from pinned memory and reads the loss back to the host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# one OpenMP thread per rank, as torchrun sets for N > 1 (the host side of a step is a handful of tiny tensor ops; an
# OpenMP team only adds wake-up latency).  Must be in the environment before torch is imported.
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

MODELS = {
    # model: (dnn, dataset, per-GPU batch (reference launch scripts), lr, preset)
    "vgg16": ("vgg16", "cifar10", 16, 0.1, "vgg16"),
    "lstman4": ("lstman4", "an4", 2, 0.001, "lstm_an4"),
    "lstm": ("lstman4", "an4", 2, 0.001, "lstm_an4"),
    "bert": ("bert_base", "wikipedia", 8, 2e-4, "bert_base"),
}


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t_mark = None

    def mark(self):
        """Start of the timed region: only samples taken after this instant are reported.  (The sampler process itself
        is started BEFORE the warm-up: nvidia-smi's NVML initialisation takes a driver-wide lock for ~0.2 s, which must not
        land inside the timed steps.)"""
        import datetime
        self.t_mark = datetime.datetime.now()

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        import datetime
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            if self.t_mark is not None:
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                    if ts < self.t_mark:
                        continue
                except ValueError:
                    pass
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--model", type=str, default=os.environ.get("OKTOPK_BENCH_MODEL", "vgg16"), choices=sorted(MODELS))
    p.add_argument("--density", type=float, default=0.001)
    p.add_argument("--compressor", type=str, default="oktopk")
    p.add_argument("--batch-size", type=int, default=None, help="per-GPU batch (default: the reference launch script's)")
    p.add_argument("--seq-len", type=int, default=128)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--backend", type=str, default=None, help="'dist' = the same scheme on NCCL collectives + torch ops (strong baseline)")
    p.add_argument("--bucket-elems", type=int, default=None)
    p.add_argument("--no-graph", action="store_true", help="disable whole-step CUDA graphs (eager launches)")
    p.add_argument("--dense-warmup", type=int, default=None,
                   help="dense warm-up iterations before the sparse phase (default: the workload preset's, as the reference)")
    p.add_argument("--no-extra", action="store_true",
                   help="flagship only: do not append the LSTM-AN4 / BERT sub-results (BASELINE configs #3, #4)")
    p.add_argument("--extra-steps", type=int, default=10)
    p.add_argument("--trace", type=str, default=None,
                   help="directory: every rank dumps the per-call device trace ring (counts, thresholds, phase times) there")
    p.add_argument("--slot-factor", type=float, default=None, help="bounded slots (default: lossless layout)")
    p.add_argument("--comm-ctas", type=int, default=None, help="CTAs of the persistent communication kernels (default: 1 per SM)")
    return p.parse_args(argv)


# -------------------------------------------------------------------------------------------------------------
# Synthetic batches.  BOTH arms train on the same stream: the formulas below are the ones of baseline/ref_runner.py
# (seed 1234 + 977 i + rank; N(0,1) images labelled by a fixed random linear teacher; AN4-shaped utterances made of
# per-character spectral templates; Wikipedia-shaped masked-LM features), so that the losses the two arms print at
# the same step are comparable.
# -------------------------------------------------------------------------------------------------------------
def make_batch(model: str, i: int, rank: int, bs: int, seq: int):
    import torch
    if model == "vgg16":
        tw = torch.randn(3 * 32 * 32, 10, generator=torch.Generator().manual_seed(4242))
        g = torch.Generator().manual_seed(1234 + 977 * i + rank)
        x = torch.randn(bs, 3, 32, 32, generator=g)
        return (x, (x.flatten(1) @ tw).argmax(1))
    if model in ("lstman4", "lstm"):
        tpl = torch.randn(29, 161, generator=torch.Generator().manual_seed(4243))
        g = torch.Generator().manual_seed(1234 + 977 * i + rank)
        tl = int(torch.randint(8, 34, (1,), generator=g))
        T = 12 * tl
        tg = torch.randint(1, 29, (bs * tl,), generator=g, dtype=torch.int32)
        base = tpl[tg.long()].view(bs, tl, 161).repeat_interleave(12, dim=1).transpose(1, 2)
        x = (base + 0.5 * torch.randn(base.shape, generator=g)).unsqueeze(1).contiguous()
        # our trainer's AN4 batch format: (inputs, concatenated targets, input length fractions, target sizes)
        return (x, tg, torch.ones(bs), torch.full((bs,), tl, dtype=torch.int32))
    vocab = 30522
    g = torch.Generator().manual_seed(4321 + 977 * i + rank)
    ids = torch.randint(1000, vocab, (bs, seq), generator=g)
    seg = (torch.arange(seq).unsqueeze(0) >= torch.randint(seq // 4, 3 * seq // 4, (bs, 1), generator=g)).long()
    lens = torch.randint(seq // 2, seq + 1, (bs, 1), generator=g)
    mask = (torch.arange(seq).unsqueeze(0) < lens).long()
    sel = (torch.rand(bs, seq, generator=g) < 0.15) & mask.bool()
    labels = torch.where(sel, ids, torch.full_like(ids, -1))
    ids = torch.where(sel, torch.full_like(ids, 103), ids) * mask
    nxt = torch.randint(0, 2, (bs,), generator=g)
    return (ids, seg, mask, labels, nxt)          # our trainer's BERT batch order


def canonical_config(model, dnn, dataset, bs, world, seq, compressor, density, n_params, dense_warm, phase):
    """The SAME keys (and, on the same workload, the same values) in both arms."""
    return {"model": dnn, "dataset_shape": dataset, "global_batch": bs * world, "per_gpu_batch": bs,
            "seq_len": seq if model == "bert" else None, "parallelism": "dp%d" % world, "compressor": compressor,
            "density": density, "params": n_params, "dense_warmup_steps_untimed": dense_warm, "timed_phase": phase,
            "l2": "no explicit flush: params+grads+residual+momentum working set %.0f MB vs 126 MB L2" % (n_params * 20 / 1e6),
            "math": "fp32 storage and accumulation; torch defaults (cuDNN conv TF32 allowed, fp32 matmul)"}


SPARSE_PHASE = "sparse (after the workload's hard-coded dense warm-up, as in the reference)"


def _loss_file(model: str, n: int) -> str:
    return os.path.join(tempfile.gettempdir(), "oktopk_bench_refloss_%s_n%d.json" % (model, n))


def run_model(args, model: str, w, steps: int, warmup: int, with_clocks: bool, do_e2e: bool) -> dict:
    """Benchmark ONE workload on the already-initialised world; returns the result dict (rank 0) or None."""
    import math
    import torch
    import torch.distributed as dist
    import oktopk_b200 as okt
    from oktopk_b200.ops import ext
    from oktopk_b200.train.trainer import Trainer

    dnn, dataset, bs0, lr, preset = MODELS[model]
    bs = args.batch_size or bs0
    # Protocol = the reference's: its hard-coded dense warm-up iterations (512 VGG / 128 LSTM / 0 BERT) run
    # first, UNTIMED, then the sparse phase is measured (the reference arm does exactly the same).
    over = dict(density=args.density)
    if args.dense_warmup is not None:
        over["warmup_iters"] = args.dense_warmup
    if args.bucket_elems:
        over["bucket_elems"] = args.bucket_elems
    if args.slot_factor is not None:
        over["slot_factor"] = over["gather_factor"] = args.slot_factor
    if args.comm_ctas is not None:
        over["comm_ctas"] = args.comm_ctas
    cfg = okt.preset(preset, **over)
    tr = Trainer(dnn=dnn, dataset=dataset, batch_size=bs, lr=lr, compressor=args.compressor, density=args.density,
                 compression=args.compressor != "none", cfg=cfg, world=w, seq_len=args.seq_len, backend=args.backend,
                 t_total=100000, warmup=0.1, cuda_graph=not args.no_graph)
    dev = tr.device
    if model == "vgg16":
        # the reference arm trains with a constant lr (its DLTrainer's multi-worker lr warm-up is not on the bench path):
        # do the same here so that the two loss curves are comparable
        tr.adjust_learning_rate = lambda: lr
        for g in tr.optimizer.param_groups:
            g["lr"] = lr

    def sync_all():
        torch.cuda.synchronize()
        if w.size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- arm 1: device-resident batches (the same 4 as the reference arm), device-timed --------------------------
    pool = [tuple(t.to(dev) for t in make_batch(model, i, w.rank, bs, args.seq_len)) for i in range(4)]
    torch.cuda.synchronize()

    def step_resident(i):
        tr.net.train()
        if tr.graphed is not None and tr.graphed.enabled:
            tr.adjust_learning_rate()
            loss = tr.graphed.step(pool[i % len(pool)])
            tr._bookkeep_iter()
            return loss
        tr.optimizer.zero_grad()
        loss, _ = tr._forward_loss(pool[i % len(pool)])
        loss.backward()
        tr.update_model()
        return loss

    dense_warm = int(cfg.warmup_iters) if args.compressor != "none" else 0
    sampler = ClockSampler(torch.cuda.current_device()) if with_clocks else None
    if sampler is not None and w.rank == 0:
        sampler.start()
    it = 0
    for _ in range(dense_warm + warmup):
        step_resident(it)
        it += 1
    sync_all()
    # Python's cyclic garbage collector must not fire inside a 25 ms timed window: a generation-2 pass over this process'
    # objects (model, graphs, thousands of tensors) takes tens of milliseconds and lands in whichever call happens to
    # allocate (measured: e2e 1.21 vs 3.0 ms/step between otherwise identical runs).  Collect now, switch it off for the
    # timed regions, back on afterwards.  (The library side: GraphedTrainStep freezes the long-lived objects, see there.)
    import gc
    gc.collect()
    gc.disable()
    if sampler is not None:
        sampler.mark()
    launches0 = ext.LAUNCH_COUNT["total"]
    marks = sorted({0, steps // 2, steps - 1})
    kept = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step_resident(it)
        it += 1
        if i in marks:
            kept[i] = loss.detach().clone()          # 3 tiny device copies; read after the timed region
    e1.record()
    sync_all()
    gc.enable()
    clocks = sampler.stop() if (sampler is not None and w.rank == 0) else None
    launches = ext.LAUNCH_COUNT["total"] - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if w.size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = bs * w.size * steps / (ms_total * 1e-3)
    losses = {"step%d" % (dense_warm + warmup + k): float(v) for k, v in sorted(kept.items())}
    stats_timed = tr.optimizer.comm_stats()

    # ---- a full schedule period outside the K steps: 64 more steps contain the 1-in-32 exact-threshold iterations and a
    #      region re-partition, which a 20-step window usually misses -> amortised step time, reported separately
    amort = None
    if args.compressor == "oktopk" and steps < 64:
        period = 64
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(period):
            step_resident(it)
            it += 1
        a1.record()
        sync_all()
        am = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
        if w.size > 1:
            dist.all_reduce(am, op=dist.ReduceOp.MAX)
        amort = {"steps": period, "ms_per_step": float(am) / period,
                 "note": "one full schedule period (exact-threshold + re-partition iterations included), device-timed, max over ranks"}

    # ---- arm 2: end to end through the public API ---------------------------------------------------------------
    e2e = None
    if do_e2e:
        for _ in range(max(3, warmup // 2)):
            tr.train_step()
            tr.record_loss()
        tr.flush_losses()
        sync_all()
        h2d0 = tr.prefetch.h2d_bytes
        for k in tr.host_us:
            tr.host_us[k] = 0.0
        gc.collect()
        gc.disable()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        for _ in range(steps):
            tr.train_step()                 # H2D of this step's batch from pinned memory happens inside (Prefetcher)
            tr.record_loss()                # D2H of the step's loss into pinned host memory (asynchronous, 4 bytes)
        e2e_losses = tr.flush_losses()      # waits for every loss copy: all K results are on the host inside the timed region
        f1.record()
        sync_all()
        gc.enable()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([max(f0.elapsed_time(f1), 0.0)], device=dev, dtype=torch.float64)
        if w.size > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"value": bs * w.size * steps / (float(ms2) * 1e-3), "unit": "samples/s",
               "h2d_bytes_per_step": (tr.prefetch.h2d_bytes - h2d0) // steps, "d2h_bytes_per_step": 4,
               "ms_per_step": float(ms2) / steps, "wall_ms_per_step": wall * 1e3 / steps,
               "losses_read": len(e2e_losses),
               "host_us_per_step": {k: round(v / max(tr.host_us["steps"], 1), 1) for k, v in tr.host_us.items() if k != "steps"},
               "api": "Trainer.train_step() + Trainer.record_loss(): DataLoader -> pinned staging ring -> H2D -> step; loss D2H per step"}
    stats = tr.optimizer.comm_stats()
    n_params = sum(p.numel() for p in tr.net.parameters())
    if args.trace:
        os.makedirs(args.trace, exist_ok=True)
        with open(os.path.join(args.trace, "trace_%s_n%d_rank%d.json" % (model, w.size, w.rank)), "w") as f:
            json.dump(tr.optimizer._allreducer.trace(), f)
    final = losses[sorted(losses, key=lambda k: int(k[4:]))[-1]] if losses else float("nan")
    # loss parity against the reference arm (it leaves its losses in a side file when it ran on this box before us)
    ref_loss, check = None, "no reference loss on this box"
    try:
        with open(_loss_file(model, w.size)) as f:
            ref_loss = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    init_level = {"vgg16": math.log(10.0), "bert": math.log(30522.0) + math.log(2.0)}.get(model)
    if not math.isfinite(final):
        check = "FAILED: non-finite loss"
    elif ref_loss and ref_loss.get("final") is not None and math.isfinite(ref_loss["final"]):
        rf = ref_loss["final"]
        if final > 1.5 * rf and init_level is not None and final > 1.5 * init_level:
            check = "FAILED: loss %.4f > 1.5x reference %.4f and above 1.5x the untrained level" % (final, rf)
        else:
            check = "ok: %.4f vs reference %.4f (ratio %.3f)" % (final, rf, final / max(rf, 1e-12))
    out = {
        "metric": "train_samples_per_sec_%s_oktopk_density%g" % (model, args.density),
        "value": value, "unit": "samples/s", "n_gpus": w.size, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "impl": "ours",
        "config": canonical_config(model, dnn, dataset, bs, w.size, args.seq_len, args.compressor, args.density, n_params,
                                   dense_warm, SPARSE_PHASE),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "amortised": amort,
        "loss": losses, "final_loss": final, "reference_loss": ref_loss, "loss_check": check,
        "arm_details": {"buckets": len(stats), "bucket_elems": cfg.bucket_elems, "comm_ctas": cfg.comm_ctas, "backend": args.backend or "cuda (fused peer-memory kernels)",
                        "channels_last": bool(getattr(tr, "channels_last", False)),
                        "fused_bn_relu_maxpool": bool(getattr(tr.net, "fuse", False)) and bool(getattr(tr, "channels_last", False)),
                        "cuda_graph": (None if tr.graphed is None else {"enabled": tr.graphed.enabled,
                                                                         "graphs": len(tr.graphed.graphs),
                                                                         "why_disabled": tr.graphed.why_disabled})},
        "comm": {k: {kk: v[kk] for kk in ("mode", "local_count", "global_count", "volume_elems", "overflow_send",
                                          "overflow_gather", "cum_overflow_send", "cum_overflow_gather", "cum_redo",
                                          "lossless", "nvls", "fault", "phase_us") if kk in v}
                 for k, v in stats_timed.items()},
    }
    tr.close()
    del tr
    torch.cuda.empty_cache()
    return out if w.rank == 0 else None


def run_ours(args) -> dict:
    import torch
    import oktopk_b200 as okt
    from oktopk_b200.ops import ext

    w = okt.init()
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local % torch.cuda.device_count())
    ext.require()
    out = run_model(args, args.model, w, args.steps, args.warmup, with_clocks=True, do_e2e=not args.no_e2e)
    if args.model == "vgg16" and not args.no_extra and os.environ.get("OKTOPK_BENCH_EXTRA", "1") == "1":
        extra = {}
        for m in ("lstman4", "bert"):
            try:
                sub = run_model(args, m, w, max(3, min(args.steps, args.extra_steps)), 3, with_clocks=False, do_e2e=False)
                if sub is not None:
                    extra[m] = {k: sub[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "config", "gpu_launches",
                                                    "final_loss", "reference_loss", "loss_check", "comm")}
            except Exception as e:  # noqa: BLE001 - a sub-result must never take the flagship number down
                extra[m] = {"unavailable": repr(e)[:300]}
        if out is not None:
            out["extra_models"] = extra
    okt.shutdown()
    return out


def main(argv=None) -> int:
    args = parse_args(argv)
    if args.impl == "reference":
        from baseline.ref_runner import run_reference_with_extras
        out = run_reference_with_extras(args, MODELS)
    else:
        out = run_ours(args)
    rc = 0
    if out is not None:
        print(json.dumps(out))
        sys.stdout.flush()
        if str(out.get("loss_check", "")).startswith("FAILED"):
            rc = 3
    return rc


if __name__ == "__main__":
    sys.exit(main())
