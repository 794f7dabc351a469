#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (read on the CPU box; needs only the `ncu` CLI).

  python tools/ncu_summary.py launches <launches.csv> <steps>            per-kernel launch list (durations + share of the step)
  python tools/ncu_summary.py full <report.ncu-rep>                      key metrics per captured launch (one --set full capture)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__waves_per_multiprocessor", "waves"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def launches(path, steps):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:64], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = max(a[0] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches (gpu__time_duration.sum, us); share = kernel class time / all captured")
    print(f"{'kernel':64s} {'launches':>8s} {'avg_us':>9s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:64s} {a[0]:8d} {a[1] / a[0]:9.2f} {a[1] / tot * 100:6.1f}%")
    if steps:
        print(f"# sum over one step ~ {tot / float(steps):.1f} us ({steps} steps captured)")
    else:
        print(f"# most frequent kernel launched {n} times")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none, one block per captured launch")
    for row in rows[2:]:
        print(f"--- {row[hdr.index('Kernel Name')][:110]}")
        for key, short in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {short:24s} {row[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
