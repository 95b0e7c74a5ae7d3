#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: device time per kernel family (shares).
    python scripts/launch_summary.py gpurun_out/launches_vgg_r2.csv [steps] > profiles/launches_vgg_r2.md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        k = r[ki].split("(")[0][:100]
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print("# Launch list: %s\n" % path.split("/")[-1])
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` over %d launches of eager (non-graph) steps; serialised, "
          "cold caches: compare SHARES, not absolutes.  Total %.0f us%s.\n" % (sum(cnt.values()), T,
          (" = %.0f us per step over ~%.1f steps" % (T / steps, steps)) if steps else ""))
    print("| device time (us) | share | launches | kernel |")
    print("|---|---|---|---|")
    for k, v in tot.most_common(40):
        print("| %.1f | %.1f %% | %d | `%s` |" % (v, 100 * v / T, cnt[k], k))
    mine = sum(v for k, v in tot.items() if "okt::" in k)
    print("\nKernels of this repo (`okt::*`): %.1f %% of the device time." % (100 * mine / T))


if __name__ == "__main__":
    main()
