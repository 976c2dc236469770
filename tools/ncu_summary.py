#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text/JSON files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/launches_X.md
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/prof_X.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def short(name):
    m = re.search(r"(\w+_kernel\w*)", name)
    if m and ("unnamed" in name or "anonymous" in name):
        return m.group(1)
    return re.sub(r"\(.*", "", name)[:90]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
        k = short(row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of {src}\n\n")
        f.write("(`--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised:"
                " compare SHARES)\n\n")
        f.write(f"total {total:.3f} ms over {sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | ms | share |\n|---|---|---|---|\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / total:.1f}% |\n")


def full(src, dst):
    raw = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[hdr.index('Kernel Name')])}` (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
