#!/usr/bin/env python
"""Turn the raw gpurun_out/ artefacts of profiles/run_profile.sh into the tracked summaries
under profiles/: the launch list (per-kernel totals + the dominant kernel's share of a step)
and the key ncu counters of the dominant kernel.

    python profiles/summarize.py <tag> [kernel-regex]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
COUNTERS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
]


def launch_table(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    seq = []
    for r in data:
        name = re.sub(r"^void ", "", r[ki]).split("(")[0].replace("<unnamed>::", "")
        t = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        seq.append((name, t))
    return agg, seq


def ncu_counters(tag, stem):
    rep = os.path.join(OUT, "%s_%s.ncu-rep" % (stem, tag))
    if not os.path.exists(rep):
        return None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H = rows[0]
    out = {}
    for c in COUNTERS:
        if c in H:
            i = H.index(c)
            out[c] = {"unit": rows[1][i], "per_launch": [r[i] for r in rows[2:]]}
    ki = H.index("Kernel Name")
    out["kernel"] = rows[2][ki]
    return out


def main():
    tag = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "score_pool_kernel")
    stem = sys.argv[3] if len(sys.argv) > 3 else "prof_score"
    agg, seq = launch_table(tag)
    total = sum(a[1] for a in agg.values())
    lines = ["# launch list `%s` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)" % tag,
             "", "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, c, t, t / c, 100 * t / total))
    # share of the dominant kernel inside one step: from one dominant launch to the next
    idx = [i for i, (n, _) in enumerate(seq) if pat.search(n)]
    if len(idx) >= 3:
        a, b = idx[-3], idx[-2]
        step = seq[a:b]
        dom = sum(t for n, t in step if pat.search(n))
        tot = sum(t for _, t in step)
        lines += ["", "One steady-state step (between two consecutive launches of the dominant kernel): "
                  + ", ".join("`%s` %.1f us" % (n, t) for n, t in step),
                  "", "Dominant kernel share of the step: **%.1f%%** (%.1f of %.1f us)" % (100 * dom / tot, dom, tot)]
    cnt = ncu_counters(tag, stem)
    if cnt:
        lines += ["", "# ncu --set full, kernel `%s`" % cnt.pop("kernel"), "", "| counter | unit | value(s) |", "|---|---|---|"]
        for k, v in cnt.items():
            lines.append("| %s | %s | %s |" % (k, v["unit"], ", ".join(v["per_launch"])))
    dst = os.path.join(ROOT, "profiles", "%s_summary.md" % tag)
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
