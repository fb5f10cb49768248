#!/usr/bin/env python
"""Per-CTA timeline of splat_class_kernel (needs a library built with -DPBRT_CLASS_TRACE, selected by PBRT_B200_LIB).
Prints how CTA durations and SM finishing times spread: the grid is one wave, so the slowest SM sets the kernel time."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pbrt_b200 as pb
from pbrt_b200 import _lib, synth
from pbrt_b200 import dist as pdist
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
pb.init(0)
W, H = wl["res"]; spp = wl["spp"]
cls = {"gaussian": pb.GaussianFilter, "mitchell": pb.MitchellFilter, "lanczos": pb.LanczosSincFilter}[wl["filter"]]
filt = cls(wl["radius"], wl["p0"], wl["p1"]) if wl["filter"] == "mitchell" else cls(wl["radius"], wl["p0"])
film = pb.Film.new([W, H], [[0, 0], [1, 1]], filt, 35.0, "t.pfm", 1.0, float("inf"))
sb = film.cropped_pixel_bounds
xy, rgbw, n = synth.samples(sb.as4(), spp, seed=1, index_bounds=sb.as4())
sbl = [[sb.p_min.x, sb.p_min.y], [sb.p_max.x, sb.p_max.y]]
for _ in range(5):
    film.add_samples_tile(sbl, spp, xy, rgbw, pb.SPLAT_EXACT)
torch.cuda.synchronize()
f = _lib.lib.pbrt_b200_debug_cta_trace
f.restype = C.c_int
buf = np.zeros((65536, 4), np.uint64); cols = C.c_int(0)
n = f(buf.ctypes.data_as(C.c_void_p), 65536, C.byref(cols))
t = buf[:n].astype(np.int64); cols = cols.value
t0 = t[:, 1].min()
start, end, sm, rows = t[:, 1] - t0, t[:, 2] - t0, t[:, 0], t[:, 3]
dur = end - start
print(f"{n} CTAs, {cols} strips x {n // cols} row segments; kernel span {end.max() / 1e3:.1f} us")
print(f"CTA start: max {start.max() / 1e3:.1f} us; duration us: min {dur.min() / 1e3:.1f} p10 {np.percentile(dur, 10) / 1e3:.1f} "
      f"median {np.median(dur) / 1e3:.1f} p90 {np.percentile(dur, 90) / 1e3:.1f} max {dur.max() / 1e3:.1f}")
smend = {}
smn = {}
for s, e in zip(sm, end):
    smend[s] = max(smend.get(s, 0), e); smn[s] = smn.get(s, 0) + 1
ends = np.array(sorted(smend.values()))
print(f"SMs {len(ends)}: finishing time us: min {ends.min() / 1e3:.1f} p10 {np.percentile(ends, 10) / 1e3:.1f} median {np.median(ends) / 1e3:.1f} "
      f"p90 {np.percentile(ends, 90) / 1e3:.1f} max {ends.max() / 1e3:.1f}; mean/max {ends.mean() / ends.max():.3f}")
print("CTAs per SM histogram:", np.bincount(list(smn.values())))
seg = np.arange(n) // cols; strip = np.arange(n) % cols
print("mean duration by strip (us):", " ".join(f"{dur[strip == c].mean() / 1e3:.0f}" for c in range(cols)))
print("mean duration by row segment (us):", " ".join(f"{dur[seg == r].mean() / 1e3:.0f}" for r in range(n // cols)))
by_cnt = {}
for s in smend: by_cnt.setdefault(smn[s], []).append(smend[s])
for k in sorted(by_cnt): print(f"SMs with {k} CTAs: {len(by_cnt[k])}, mean finish {np.mean(by_cnt[k]) / 1e3:.1f} us")
# duration vs SM id (die / GPC effects)
order = np.argsort(list(smend.keys()))
ids = np.array(list(smend.keys()))[order]; fe = np.array(list(smend.values()))[order]
print("finish by SM id (us, groups of 8):", " ".join(f"{fe[i:i + 8].mean() / 1e3:.0f}" for i in range(0, len(fe), 8)))
# the slowest and fastest SMs: their CTAs as (strip, segment, rows, duration us)
bysm = {}
for i in range(n): bysm.setdefault(int(sm[i]), []).append((int(strip[i]), int(seg[i]), int(rows[i]), round(dur[i] / 1e3)))
worst = sorted(smend, key=lambda s: -smend[s])
for s in worst[:6] + worst[-3:]:
    print(f"SM {s:3d} finish {smend[s] / 1e3:.0f} us:", bysm[int(s)])
