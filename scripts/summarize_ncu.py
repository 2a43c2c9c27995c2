#!/usr/bin/env python
"""Turn ncu exports into the summaries kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python scripts/summarize_ncu.py raw raw.csv "title" >> profiles/rNN_ncu_summary.md
    python scripts/summarize_ncu.py launches launches.csv > profiles/rNN_launches.md
"""
import collections
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]


def raw(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"## {title}\n")
    for r in rows[2:]:
        print(f"### {r[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                print(f"- {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}")
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and r[i]]
        top = ", ".join(f"{h.split('stalled_')[1].split('_per')[0]} {v:.2f}" for v, h in sorted(st, reverse=True)[:5])
        print(f"- top stalls (warps per issue): {top}\n")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a[0] += 1
        a[1] += float(r[-1])
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share | grid | block |\n|---|---|---|---|---|---|---|")
    for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.2f} | {100 * ns / total:.1f}% | {grid} | {block} |")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[2])
    else:
        launches(sys.argv[2])
