#!/usr/bin/env python
"""Summarise an ncu report of one kernel: headline metrics, instruction mix, stall reasons, hot blocks.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [samples_per_launch]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else None


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, vals = raw[0], raw[2]
m = dict(zip(hdr, vals))
print("kernel:", m.get("Kernel Name"))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct", "sm__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
units = dict(zip(hdr, raw[1]))
for k in keys:
    if k in m:
        print(f"  {k:72s} {m[k]:>16s} {units.get(k, '')}")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv"))))
h, data = src[1], src[2:]
ix = {n: i for i, n in enumerate(h)}


def g(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(g(r, "Instructions Executed") for r in data)
tots = sum(g(r, "# Samples") for r in data)
per = (lambda n: f"{n * 32 / nsamp:7.1f}/sample") if nsamp else (lambda n: "")
print(f"warp instructions {tot:.4e}" + (f" = {tot * 32 / nsamp:.1f} thread-instructions per sample" if nsamp else ""))
c, cs = Counter(), Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    c[op] += g(r, "Instructions Executed")
    cs[op] += g(r, "# Samples")
print("instruction mix (executed share, per sample, share of stall samples):")
for op, n in c.most_common(24):
    print(f"  {op:22s} {n / tot * 100:6.2f}% {per(n)}  {cs[op] / max(tots, 1) * 100:6.2f}%")
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print("stall samples:", ", ".join(f"{n[6:]}={v / max(tots, 1) * 100:.1f}%" for v, n in sorted(((sum(g(r, n) for r in data), n) for n in stalls), reverse=True)[:9]))
blocks, cur = [], None
for i, r in enumerate(data):
    e = g(r, "Instructions Executed")
    if cur and abs(cur[2] - e) < 1e-9:
        cur[1] = i
        cur[3] += g(r, "# Samples")
    else:
        cur = [i, i, e, g(r, "# Samples")]
        blocks.append(cur)
print("hot straight-line blocks (>1.5% of executed instructions):")
for b in blocks:
    n = b[1] - b[0] + 1
    share = b[2] * n / tot * 100
    if share > 1.5:
        print(f"  sass {b[0]:4d}-{b[1]:4d} n={n:3d} exec/inst={b[2]:.3e} share={share:5.1f}% stall-samples={b[3] / max(tots, 1) * 100:4.1f}%  {data[b[0]][ix['Source']][:44]}")
