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
    p.add_argument("--backend", type=str, default=None)
    p.add_argument("--bucket-elems", type=int, default=None)
    p.add_argument("--no-graph", action="store_true", help="disable whole-step CUDA graphs (eager launches)")
    p.add_argument("--dense-warmup", type=int, default=None,
                   help="dense warm-up iterations before the sparse phase (default: the workload preset's, as the reference)")
    return p.parse_args(argv)


def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist
    import oktopk_b200 as okt
    from oktopk_b200.ops import ext
    from oktopk_b200.train.trainer import Trainer

    w = okt.init()
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local % torch.cuda.device_count())
    ext.require()
    dnn, dataset, bs0, lr, preset = MODELS[args.model]
    bs = args.batch_size or bs0
    # Protocol = the reference's: its hard-coded dense warm-up iterations (512 VGG / 128 LSTM / 0 BERT) run
    # first, UNTIMED, then the sparse phase is measured (the reference arm does exactly the same).
    over = dict(density=args.density)
    if args.dense_warmup is not None:
        over["warmup_iters"] = args.dense_warmup
    if args.bucket_elems:
        over["bucket_elems"] = args.bucket_elems
    cfg = okt.preset(preset, **over)
    tr = Trainer(dnn=dnn, dataset=dataset, batch_size=bs, lr=lr, compressor=args.compressor, density=args.density,
                 compression=args.compressor != "none", cfg=cfg, world=w, seq_len=args.seq_len, backend=args.backend,
                 t_total=100000, warmup=0.1, cuda_graph=not args.no_graph)
    dev = tr.device

    def sync_all():
        torch.cuda.synchronize()
        if w.size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- arm 1: device-resident batches, device-timed ---------------------------------------------
    pool = [tr.prefetch.next() for _ in range(4)]          # a few distinct resident batches
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
    sampler = ClockSampler(torch.cuda.current_device())
    if w.rank == 0:
        sampler.start()
    for i in range(dense_warm + args.warmup):
        step_resident(i)
    sync_all()
    sampler.mark()
    launches0 = ext.LAUNCH_COUNT["total"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    sync_all()
    clocks = sampler.stop() if w.rank == 0 else None
    launches = ext.LAUNCH_COUNT["total"] - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if w.size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = bs * w.size * args.steps / (ms_total * 1e-3)

    # ---- arm 2: end to end through the public API -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(max(3, args.warmup // 4)):
            tr.train_step()
            tr.last_loss()
        sync_all()
        h2d0 = tr.prefetch.h2d_bytes
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record()
        d2h = 0
        for _ in range(args.steps):
            tr.train_step()                 # H2D of this step's batch from pinned memory happens inside (Prefetcher)
            _ = tr.last_loss()              # D2H read of the step's loss
            d2h += 4
        f1.record()
        sync_all()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([max(f0.elapsed_time(f1), 0.0)], device=dev, dtype=torch.float64)
        if w.size > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"value": bs * w.size * args.steps / (float(ms2) * 1e-3), "unit": "samples/s",
               "h2d_bytes_per_step": (tr.prefetch.h2d_bytes - h2d0) // args.steps, "d2h_bytes_per_step": d2h // args.steps,
               "ms_per_step": float(ms2) / args.steps, "wall_ms_per_step": wall * 1e3 / args.steps}
    stats = tr.optimizer.comm_stats()
    n_params = sum(p.numel() for p in tr.net.parameters())
    working_set_mb = n_params * 4 * 5 / 1e6
    out = {
        "metric": "train_samples_per_sec_%s_oktopk_density%g" % (args.model, args.density),
        "value": value, "unit": "samples/s", "n_gpus": w.size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic", "impl": "ours",
        "config": {"model": dnn, "dataset_shape": dataset, "global_batch": bs * w.size, "per_gpu_batch": bs,
                   "seq_len": args.seq_len if args.model == "bert" else None, "parallelism": "dp%d" % w.size,
                   "compressor": args.compressor, "density": args.density, "params": n_params,
                   "l2": "no explicit flush: params+grads+residual+momentum working set %.0f MB vs 126 MB L2" % working_set_mb,
                   "buckets": len(stats), "dense_warmup_steps_untimed": dense_warm,
                   "timed_phase": "sparse (after the workload's dense warm-up, as in the reference)",
                   "math": "fp32 storage and accumulation; torch defaults (cuDNN conv TF32 allowed, fp32 matmul)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "cuda_graph": (None if tr.graphed is None else {"enabled": tr.graphed.enabled, "graphs": len(tr.graphed.graphs),
                                                        "why_disabled": tr.graphed.why_disabled}),
        "final_loss": tr.last_loss(),
        "comm": {k: {kk: v[kk] for kk in ("mode", "local_count", "global_count", "volume_elems", "overflow_send",
                                          "overflow_gather", "fault", "phase_us") if kk in v} for k, v in stats.items()},
    }
    tr.close()
    okt.shutdown()
    return out if w.rank == 0 else None


def main(argv=None) -> int:
    args = parse_args(argv)
    if args.impl == "reference":
        from baseline.ref_runner import run_reference
        out = run_reference(args, MODELS)
    else:
        out = run_ours(args)
    if out is not None:
        print(json.dumps(out))
        sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
