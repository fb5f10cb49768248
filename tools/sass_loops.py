#!/usr/bin/env python
"""Static loop census of a SASS listing made by tools/sass_probe.sh: every backward branch closes a loop; print each
loop's length and instruction mix (loops that contain FADD2 / FFMA are the gather bodies)."""
import re, sys
from collections import Counter
lines = [l.split(None, 1) for l in open(sys.argv[1] if len(sys.argv) > 1 else "/tmp/probe/p.sass") if l.strip()]
addr = {int(a, 16): i for i, (a, _) in enumerate(lines)}
for i, (a, ins) in enumerate(lines):
    m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)", ins)
    if not m: continue
    t = int(m.group(1), 16)
    if t in addr and addr[t] <= i:
        body = [x[1] for x in lines[addr[t]:i + 1]]
        ops = Counter()
        for b in body:
            tok = b.split()
            op = tok[1] if tok[0].startswith("@") else tok[0]
            ops[op.split(".")[0]] += 1
        if ops["FADD2"] + ops["FFMA"] + ops["FADD"] < 4: continue
        taps = ops["FADD2"] / 2 if ops["FADD2"] else ops["FFMA"] / 3
        print(f"loop {addr[t]:5d}-{i:5d} n={len(body):4d} taps={taps:5.1f} per-tap={len(body) / max(taps, 1):5.2f} | " +
              " ".join(f"{k}:{v}" for k, v in ops.most_common(9)))
