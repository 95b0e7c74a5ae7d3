#!/usr/bin/env python
"""Distil ncu exports (``--page raw --csv`` / ``--page source --csv``) into short tracked summaries under profiles/.
    python scripts/ncu_summary.py gpurun_out/okt_vgg_d001 [more prefixes...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
    ("lts__t_sectors_op_red.sum", "L2 reduction sectors"), ("lts__t_sectors_op_atom.sum", "L2 atomic sectors"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        rec = {"kernel": d.get("Kernel Name", ("?", ""))[0]}
        for k, label in KEYS:
            if k in d and d[k][0] != "":
                rec[label] = "%s %s" % d[k]
        st = []
        for h, (v, u) in d.items():
            if h.startswith(STALLS) and h.endswith("_per_warp_active.pct"):
                try:
                    st.append((float(v), h[len(STALLS):-len("_per_warp_active.pct")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        rec["top stalls (% of warp-active)"] = ", ".join("%s %.0f" % (n, v) for v, n in st[:6])
        out.append(rec)
    return out


def source_summary(path, top=14):
    rows = list(csv.reader(open(path)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    res = []
    for si, s in enumerate(secs):
        e = secs[si + 1] if si + 1 < len(secs) else len(rows)
        hdr = rows[s + 1]
        ci = {h: i for i, h in enumerate(hdr)}
        body = rows[s + 2:e]
        key = "# Samples" if "# Samples" in ci else "Warp Stall Sampling (All Samples)"
        def n(r):
            try:
                return int(r[ci[key]] or 0)
            except ValueError:
                return 0
        tot = sum(n(r) for r in body) or 1
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        lines = []
        for r in sorted(body, key=lambda r: -n(r))[:top]:
            st = sorted([(int(r[ci[h]] or 0), h[6:]) for h in stalls], reverse=True)[:2]
            lines.append("%5.1f%%  %-58s %s" % (100.0 * n(r) / tot, r[ci["Source"]].strip()[:58],
                                              " ".join("%s:%d" % (b, a) for a, b in st if a)))
        res.append({"kernel": rows[s][1], "samples": tot, "hot": lines})
    return res


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for prefix in sys.argv[1:]:
        name = os.path.basename(prefix)
        md = ["# ncu summary: %s" % name, "",
              "`ncu --set full --clock-control none --launch-count 1` (B200, one GPU, after warm-up); distilled by "
              "scripts/ncu_summary.py from the `--page raw --csv` / `--page source --csv` exports.", ""]
        if os.path.exists(prefix + ".raw.csv"):
            for rec in raw_summary(prefix + ".raw.csv"):
                md.append("## %s" % rec.pop("kernel"))
                md.append("")
                md.append("| metric | value |")
                md.append("|---|---|")
                for k, v in rec.items():
                    md.append("| %s | %s |" % (k, v))
                md.append("")
        if os.path.exists(prefix + ".source.csv"):
            for rec in source_summary(prefix + ".source.csv"):
                md.append("### hottest SASS lines by warp-stall samples (%d samples)" % rec["samples"])
                md.append("")
                md.append("```")
                md.extend(rec["hot"])
                md.append("```")
                md.append("")
        with open(os.path.join(ROOT, "profiles", "ncu_%s.md" % name), "w") as f:
            f.write("\n".join(md) + "\n")
        print("wrote profiles/ncu_%s.md" % name)


if __name__ == "__main__":
    main()
