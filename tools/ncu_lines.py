#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report (needs -lineinfo and --import-source on).

usage: python tools/ncu_lines.py report.ncu-rep samples_per_launch [min_per_sample]
"""
import csv, io, subprocess, sys
rep, nsamp = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, tot, lines = None, None, 0.0, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < 10: continue
    if r[2] == "-":  # a source line (its sass rows follow)
        try: ti = float(r[hdr["Thread Instructions Executed"]]); smp = float(r[hdr["# Samples"]])
        except ValueError: continue
        lines.append((fname, int(r[0]), r[1].strip(), ti, smp)); tot += ti
print(f"total thread instructions {tot:.4e} = {tot / nsamp:.1f} per sample")
ssum = sum(l[4] for l in lines)
for f, n, s, ti, smp in lines:
    if ti / nsamp >= thr:
        print(f"{f}:{n:4d} {ti / nsamp:7.2f}/sample {smp / ssum * 100:5.1f}% stalls | {s[:110]}")
