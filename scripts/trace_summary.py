#!/usr/bin/env python
"""Per-rank x per-phase summary of the device-side trace rings dumped by ``bench.py --trace DIR`` (one TraceRec per fused
Ok-Topk call: phase durations from globaltimer stamps, counts, thresholds).

    python scripts/trace_summary.py gpurun_out/trace8 vgg16 8 > profiles/scaling_trace_vgg16_n8.md

``wait_rs`` is the time a rank spends waiting for the slowest peer's reduce-scatter flags after it finished its own pack
pass: per step, the rank with the smallest wait arrived last, i.e. it set the pace of that step."""
import json
import statistics as S
import sys


def main():
    d, model, P = sys.argv[1], sys.argv[2], int(sys.argv[3])
    by = {}
    for r in range(P):
        t = json.load(open("%s/trace_%s_n%d_rank%d.json" % (d, model, P, r)))
        for b, recs in t.items():
            by.setdefault(b, {})[r] = {x["epoch"]: x for x in recs}
    keys = ["us_local", "us_pack", "us_wait_rs", "us_reduce", "us_gselect", "us_wait_ag", "us_final"]
    print("# Device-side phase times of the fused Ok-Topk kernel, %s, %d GPUs\n" % (model, P))
    print("Source: `bench.py --trace` (trace ring of `csrc/oktopk.cu`, globaltimer stamps by block 0), microseconds, mean over "
          "the calls every rank still holds in its ring; `last%` = share of calls in which this rank was the last to reach the "
          "reduce-scatter handshake (smallest wait).\n")
    for b, ranks in by.items():
        common = sorted(set.intersection(*[set(v) for v in ranks.values()]))
        if not common:
            continue
        last = [0] * P
        for e in common:
            w = [ranks[r][e]["us_wait_rs"] for r in range(P)]
            last[w.index(min(w))] += 1
        print("## bucket `%s` (%d calls)\n" % (b, len(common)))
        print("| rank | " + " | ".join(k[3:] for k in keys) + " | total | last% | local_count (mean) |")
        print("|---|" + "---|" * (len(keys) + 3))
        for r in range(P):
            recs = [ranks[r][e] for e in common]
            m = [S.mean(x[k] for x in recs) for k in keys]
            print("| %d | " % r + " | ".join("%.1f" % v for v in m) + " | %.1f | %.0f | %.0f |" % (
                sum(m), 100.0 * last[r] / len(common), S.mean(x["local_count"] for x in recs)))
        allw = sorted(ranks[r][e]["us_wait_rs"] for r in range(P) for e in common)
        tot = sorted(sum(ranks[r][e][k] for k in keys) for r in range(P) for e in common)
        print("\nwait_rs over all ranks and calls: mean %.1f, p50 %.1f, p90 %.1f, max %.1f us; kernel total: p50 %.1f, p90 %.1f us\n" % (
            S.mean(allw), allw[len(allw) // 2], allw[int(0.9 * len(allw))], allw[-1], tot[len(tot) // 2], tot[int(0.9 * len(tot))]))
        ov = sum(ranks[r][e]["overflow_send"] + ranks[r][e]["overflow_gather"] for r in range(P) for e in common)
        rd = sum(ranks[r][e]["redo"] for r in range(P) for e in common)
        print("overflowed entries over all ranks and calls: %d; pack passes repeated: %d\n" % (ov, rd))


if __name__ == "__main__":
    main()
