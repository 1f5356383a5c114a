#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv        # per-kernel share of a launch list
    python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep            # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

FULL = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms total (ncu-serialised, cold cache: compare SHARES)")
    print(f"{'kernel':72s} {'n':>6s} {'total_us':>10s} {'share%':>7s} {'avg_us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {v[0]:6d} {v[1] / 1e3:10.1f} {100 * v[1] / tot:7.2f} {v[1] / v[0] / 1e3:9.2f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path} (ncu --set full --clock-control none), one block per captured launch")
    for r in rows[2:]:
        for w in FULL:
            if w in idx:
                print(f"{w:72s} {r[idx[w]]} {units[idx[w]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
