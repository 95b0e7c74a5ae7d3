"""Sparse-allreduce bandwidth sweep (BASELINE.json config #5; SURVEY 6.2 "secondary").

    torchrun --nproc-per-node P --master-addr 127.0.0.1 -m oktopk_b200.bench.sweep \
        --sizes 1M,16M,128M,1G --densities 0.1,0.01,0.001 --schemes oktopk,topkSA,gtopk,dense,nccl

For every (n, density, scheme) the bucket is reduced ``--iters`` times on the fused peer-memory kernels
(``nccl`` = ``dist.all_reduce``; ``<scheme>_nccl`` = the same sparse scheme on NCCL collectives + torch ops, the strong
baseline the fused kernels are compared against), each call timed with CUDA events on the launching
stream and reported as the MAX over ranks of the median.  Reported per row:

  * ``ms``            device time of one allreduce call,
  * ``algbw_GBs``     4n / t  (dense-equivalent algorithm bandwidth),
  * ``moved_MB``      bytes this rank pulled over NVLink (idx+val pairs: 8 B each), from the device counters,
  * ``bound_MB``      the paper's 6k(P-1)/P scalars = 6k * 8 B * (P-1)/P  (reference README.md:2),
  * ``link_GBs``      moved bytes / t  against 900 GB/s per direction per GPU,
  * ``hbm_frac``      (16 n bytes / t) / measured HBM copy bandwidth  -- the roofline of the steady-state call
                      (read grad, read+write residual, write result), since at density <= 0.01 the call
                      is HBM-bound, not link-bound.

Steady state is measured: one exact-threshold call first (untimed), then threshold-reuse calls on the same
gradient with the residual reset, so every timed call selects ~k entries.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
from typing import Dict, List

import torch
import torch.distributed as dist


def _parse_size(s: str) -> int:
    s = s.strip().upper()
    mult = {"K": 1 << 10, "M": 1 << 20, "G": 1 << 30}
    if s[-1] in mult:
        return int(float(s[:-1]) * mult[s[-1]])
    return int(s)


def _hbm_gbs() -> float:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6484.3


def run(args) -> List[Dict]:
    import oktopk_b200 as okt
    from oktopk_b200.ops import ext
    from oktopk_b200.parallel.gpu_engine import CudaBucketEngine

    w = okt.init()
    P, rank = w.size, w.rank
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    ext.require()
    hbm = _hbm_gbs()
    rows = []
    for n in args.sizes:
        src = torch.empty(n, dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        src.normal_(generator=g)
        # a smooth magnitude ramp makes the balanced regions differ from uniform ones
        src.mul_(torch.linspace(0.5, 1.5, steps=1024, device=dev).repeat_interleave((n + 1023) // 1024)[:n])
        for scheme in args.schemes:
            dens_list = [None] if scheme in ("dense", "dense_p2p", "dense_nvls", "nccl") else args.densities
            if scheme.endswith("_nccl") and n > args.nccl_max_n:
                continue
            for density in dens_list:
                d = 0.001 if density is None else density
                k = max(int(n * d), 1)
                if scheme == "gtopk" and (P & (P - 1)):
                    continue
                if scheme.endswith("_nccl"):
                    # the honest strong baseline: the SAME scheme on NCCL collectives (all_to_all / all_gather handshakes,
                    # send/recv payloads) + torch ops for selection/scatter (parallel/algorithms.py, backend='dist')
                    from oktopk_b200.parallel import algorithms
                    from oktopk_b200.parallel.state import SparseState
                    base = scheme[:-5]
                    cfg = okt.OkTopkConfig(density=d, warmup_iters=0, local_recompute_interval=1 << 30,
                                           global_recompute_interval=1 << 30, repartition_interval=1 << 30, backend="dist")
                    stt = SparseState(n, P)
                    buf = src.clone()
                    times = []
                    for it in range(args.warmup + args.iters):
                        buf.copy_(src)
                        if stt.residual is not None:
                            stt.residual.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        w.barrier()
                        a.record()
                        algorithms.sparse_allreduce(base, buf, stt, cfg, w, d)
                        b.record()
                        torch.cuda.synchronize()
                        if it >= args.warmup:
                            times.append(a.elapsed_time(b))
                    moved = 4.0 * stt.last_volume_elems
                    stats = {"local_count": stt.last_local_count, "global_count": stt.last_global_count}
                    del buf, stt
                elif scheme == "nccl":
                    buf = src.clone()
                    times = []
                    for it in range(args.warmup + args.iters):
                        buf.copy_(src)
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        w.barrier()
                        a.record()
                        if P > 1:
                            dist.all_reduce(buf)
                        buf.div_(P)
                        b.record()
                        torch.cuda.synchronize()
                        if it >= args.warmup:
                            times.append(a.elapsed_time(b))
                    moved = 2 * 4 * n * (P - 1) / P
                    stats = {}
                    del buf
                else:
                    cfg = okt.OkTopkConfig(density=d, warmup_iters=0, local_recompute_interval=1 << 30,
                                           global_recompute_interval=1 << 30, repartition_interval=1 << 30,
                                           slot_factor=args.slot_factor, gather_factor=args.slot_factor,
                                           sparse=not scheme.startswith("dense"), pull_mode=args.pull,
                                           gselect_mode=args.gselect, dense_switch_density=args.dense_switch,
                                           nvls={"dense_p2p": "off", "dense_nvls": "on"}.get(scheme, "auto"),
                                           compressor=scheme if not scheme.startswith("dense") else "none")
                    eng = CudaBucketEngine(n, cfg, w, name="sweep")
                    name = "none" if scheme.startswith("dense") else scheme
                    times = []
                    for it in range(args.warmup + args.iters):
                        eng.grad.copy_(src)
                        if it > 0 and name == "oktopk":
                            eng.residual.zero_()              # same accumulator every call => ~k selected every call
                        elif name != "oktopk":
                            eng.residual.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        w.barrier()
                        a.record()
                        eng.reduce(name)
                        b.record()
                        torch.cuda.synchronize()
                        if it >= args.warmup:
                            times.append(a.elapsed_time(b))
                    stats = eng.stats() if name != "none" else {}
                    if name == "none":
                        moved = 2 * 4 * n * (P - 1) / P
                    elif name == "gtopk":
                        moved = 8.0 * 2 * k * max(P.bit_length() - 1, 0)
                    else:
                        moved = 8.0 * (stats.get("recv_total", 0) + stats.get("gather_total", 0)) * (P - 1) / max(P, 1)
                    eng.close()
                    del eng
                ms = statistics.median(times)
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                if P > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
                row = {"n": n, "P": P, "scheme": scheme, "density": density, "k": None if density is None else k,
                       "ms": ms, "algbw_GBs": 4.0 * n / (ms * 1e-3) / 1e9, "moved_MB": moved / 1e6,
                       "bound_MB": None if density is None else 6 * k * 8.0 * (P - 1) / P / 1e6,
                       "link_GBs": moved / (ms * 1e-3) / 1e9,
                       "hbm_frac": None if scheme in ("dense", "dense_p2p", "dense_nvls", "nccl") else (16.0 * n / (ms * 1e-3) / 1e9) / hbm,
                       "mode": stats.get("mode"),
                       "local_count": stats.get("local_count"), "global_count": stats.get("global_count"),
                       "overflow": (stats.get("overflow_send", 0) + stats.get("overflow_gather", 0)) if stats else None,
                       "phase_us": {k2: round(v2, 1) for k2, v2 in stats.get("phase_us", {}).items()} if stats else None}
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
        del src
        torch.cuda.empty_cache()
    okt.shutdown()
    return rows


def to_markdown(rows: List[Dict]) -> str:
    out = ["| n | P | scheme | density | ms | alg GB/s | moved MB | 6k bound MB | link GB/s | HBM-roofline frac | sel local/global |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        out.append("| %d | %d | %s | %s | %.3f | %.1f | %.2f | %s | %.1f | %s | %s/%s |" % (
            r["n"], r["P"], r["scheme"], "-" if r["density"] is None else "%g" % r["density"], r["ms"], r["algbw_GBs"],
            r["moved_MB"], "-" if r["bound_MB"] is None else "%.2f" % r["bound_MB"], r["link_GBs"],
            "-" if r["hbm_frac"] is None else "%.2f" % r["hbm_frac"], r["local_count"], r["global_count"]))
    return "\n".join(out)


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--sizes", type=lambda s: [_parse_size(x) for x in s.split(",")], default=[1 << 20, 1 << 24, 1 << 27])
    p.add_argument("--densities", type=lambda s: [float(x) for x in s.split(",")], default=[0.1, 0.01, 0.001])
    p.add_argument("--schemes", type=lambda s: s.split(","), default=["oktopk", "topkSA", "gtopk", "dense", "nccl"])
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--slot-factor", type=float, default=0.0, help="0 = lossless slot layout (default), > 0 = bounded slots")
    p.add_argument("--nccl-max-n", type=lambda s: _parse_size(s), default=1 << 28,
                   help="largest bucket for the *_nccl (torch ops) baselines: torch.topk / nonzero temporaries are several x n")
    p.add_argument("--gselect", type=str, default="auto", choices=["auto", "list", "scan"])
    p.add_argument("--dense-switch", type=float, default=0.05, help="automatic dense switch density (0 = off)")
    p.add_argument("--pull", type=str, default="tma", choices=["tma", "ldg"])
    p.add_argument("--out", type=str, default=None, help="write a markdown table here (rank 0)")
    args = p.parse_args(argv)
    rows = run(args)
    if args.out and int(os.environ.get("RANK", "0")) == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(to_markdown(rows) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
