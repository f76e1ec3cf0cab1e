#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total device time)
and, for .ncu-rep files, print the roofline-relevant raw metrics per captured launch.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv [--ours-only]
    python tools/ncu_summary.py rep gpurun_out/prof_gemm_r1.ncu-rep
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe",
        "sm__inst_executed_pipe_tmem", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum ", "sm__cycles_active.avg", "smsp__inst_executed.sum ",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__pipe_xu_cycles_active", "smsp__inst_executed_pipe_xu",
        "sm__inst_executed_pipe_xu", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct"]


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(?:dots::)?([A-Za-z0-9_]+)", name)
    base = m.group(1) if m else name[:60]
    if name.startswith("dots::"):
        t = re.search(r"<([^>]*)>", name)
        return f"dots::{base}" + (f"<{t.group(1)}>" if t else "")
    if name.startswith("at::") or "at::native" in name or "distribution" in name:
        return "torch:" + base
    return base


def launches(path, ours_only=False):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), v_us, r["Grid Size"], r["Block Size"]))
    if ours_only:
        rows = [r for r in rows if r[0].startswith("dots::")]
    agg = OrderedDict()
    for n, us, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += us
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {us / 1e3:.3f} | {us / c:.1f} | {100 * us / total:.2f}% |")
    print(f"| **total** | {sum(a[0] for a in agg.values())} | {total / 1e3:.3f} | | 100% |")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print(f"### {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k in hdr:
            if any(k.startswith(p) for p in KEYS):
                print(f"  {k} = {r[idx[k]]} {units[idx[k]]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], "--ours-only" in sys.argv)
    else:
        rep(sys.argv[2])
