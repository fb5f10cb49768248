#!/usr/bin/env python
"""Back-to-back splat launches timed with two events around the whole loop (per-step events would sit between the
kernels and switch programmatic dependent launch off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pbrt_b200 as pb
from pbrt_b200 import synth
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pb.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); pb.set_stream(stream.cuda_stream)
W, H = wl["res"]; spp = wl["spp"]
crop = [[0, 0], [1, 1]]
if len(sys.argv) > 3:  # e.g. "1920x2160:0,0.5,1,1": another film size and crop window with the workload's filter
    r, c = sys.argv[3].split(":")
    W, H = (int(v) for v in r.split("x"))
    c = [float(v) for v in c.split(",")]
    crop = [[c[0], c[1]], [c[2], c[3]]]
if len(sys.argv) > 4:
    spp = int(sys.argv[4])
cls = {"gaussian": pb.GaussianFilter, "mitchell": pb.MitchellFilter, "lanczos": pb.LanczosSincFilter}[wl["filter"]]
filt = cls(wl["radius"], wl["p0"], wl["p1"]) if wl["filter"] == "mitchell" else cls(wl["radius"], wl["p0"])
film = pb.Film.new([W, H], crop, filt, 35.0, "t.pfm", 1.0, float("inf"))
sb = film.cropped_pixel_bounds
xy, rgbw, n = synth.samples(sb.as4(), spp, seed=1, index_bounds=sb.as4())
sbl = [[sb.p_min.x, sb.p_min.y], [sb.p_max.x, sb.p_max.y]]
for _ in range(5): film.add_samples_tile(sbl, spp, xy, rgbw, pb.SPLAT_EXACT)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): film.add_samples_tile(sbl, spp, xy, rgbw, pb.SPLAT_EXACT)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{ms:.4f} ms/launch  {n / ms / 1e-3:.4g} samples/s")
film.check()
