#!/usr/bin/env python
"""bench.py — filtered samples/s into the Film on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU path on the box's host cores

A step is one pass of the hot path over one batch of synthetic samples: get_film_tile ->
add_sample for every sample -> merge_film_tile, as ONE kernel launch per GPU.

Workload at N=1: BASELINE.json configs[1] — 1920x1080 film, 16 spp stratified samples (SURVEY.md
App. C, PCG32, seed 1), Gaussian filter radius 2 (alpha 2).  At N>1 the film is sharded by rows,
one process per GPU, and the job is weak-scaled: every rank owns a 1920x1080 row block of a
1920x(1080*N) film, so per-GPU work is fixed.  Samples within h = 2 rows of a shard edge are
processed by both neighbours (no collective on the data path); the final frame is assembled by
one NCCL all-gather of the resolved rows, timed separately (assemble_ms).

`value`   : samples/s with the sample streams resident in HBM (timed with CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks).
`e2e`     : the same through the public API with HOST buffers: pinned host samples are copied to
            the device every step and the resolved RGB frame is read back.
`roofline`: algorithmic bytes (24 B/sample + 32 B per film pixel per pass) / kernel duration
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline`: the C restatement of the reference's CPU path (oracle/, "port": the reference is
            Rust and cannot be built here), one thread, on a bounded band of the same workload.

The reference has no add_sample and only a box filter: the splat path is an
EXTENSION with no reference parity; its checker is the oracle's restatement of pbrt-v3.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # name: (width, height, spp per pass, filter name, radius, p0, p1, passes per step)
    "c2": dict(res=(1920, 1080), spp=16, filter="gaussian", radius=(2.0, 2.0), p0=2.0, p1=0.0,
               label="1920x1080 film, 16 spp stratified, Gaussian r=2 (BASELINE configs[1])"),
    "c3": dict(res=(3840, 2160), spp=16, filter="mitchell", radius=(2.0, 2.0), p0=1 / 3, p1=1 / 3,
               label="3840x2160 film, Mitchell r=2, one 16-spp pass of the 64 spp (BASELINE configs[2])"),
    "c5": dict(res=(7680, 4320), spp=16, filter="lanczos", radius=(4.0, 4.0), p0=3.0, p1=0.0,
               label="7680x4320 film, Lanczos-sinc r=4, one 16-spp pass of the 256 spp (BASELINE configs[4])"),
    "c3_spp64": dict(res=(3840, 2160), spp=64, filter="mitchell", radius=(2.0, 2.0), p0=1 / 3, p1=1 / 3,
                     label="3840x2160 film, Mitchell r=2, all 64 spp in one pixel-major pass (BASELINE configs[2])"),
    "c2_spp8": dict(res=(1920, 1080), spp=8, filter="gaussian", radius=(2.0, 2.0), p0=2.0, p1=0.0,
                    label="1920x1080 film, 8 spp per pass, Gaussian r=2 (occupancy experiment, not a BASELINE config)"),
    "c2_spp4": dict(res=(1920, 1080), spp=4, filter="gaussian", radius=(2.0, 2.0), p0=2.0, p1=0.0,
                    label="1920x1080 film, 4 spp per pass, Gaussian r=2 (occupancy experiment, not a BASELINE config)"),
    "c1": dict(res=(64, 64), spp=4, filter="gaussian", radius=(2.0, 2.0), p0=2.0, p1=0.0,
               label="64x64 film, 4 spp (BASELINE configs[0])"),
}
FILTER_KIND = {"box": 0, "triangle": 1, "gaussian": 2, "mitchell": 3, "lanczos": 4}


def measured_traffic(workload: str, mode: str):
    """DRAM bytes per launch of the splat kernel from the committed ncu capture (profiles/traffic.json)."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(f"{workload}:{mode}")
        except Exception:
            pass
    return None


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (via NVML)."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------- CPU arm

class CpuArm:
    """The oracle's splat pass over a band of the workload: `band_rows` pixel rows of the film."""

    def __init__(self, wl: dict, band_rows: int):
        import oracle
        from oracle import OracleFilm

        o = oracle.load()
        W, H = wl["res"]
        band_rows = min(band_rows, H)
        table = oracle.filter_table(o, FILTER_KIND[wl["filter"]], wl["radius"], wl["p0"], wl["p1"])
        y0 = (H - band_rows) // 2
        y1 = y0 + band_rows
        # the band is a Film whose crop window is those rows (src/core/film.rs:92-101 supports this natively)
        self.film = OracleFilm(o, (W, H), [0.0, y0 / H, 1.0, y1 / H], wl["radius"], table)
        assert self.film.cropped() == (0, y0, W, y1), self.film.cropped()
        self.sb = (0, y0, W, y1)
        self.spp = wl["spp"]
        self.xy, self.rgbw = oracle.synth_samples(o, self.sb, self.spp, 1)
        self.n = len(self.xy)
        self.desc = f"{W}x{band_rows} pixel band of the workload ({self.n} samples per step)"

    def step(self, threads: int) -> float:
        t0 = time.perf_counter()
        self.film.add_samples_pass(self.sb, self.spp, self.xy, self.rgbw, threads=threads)
        return time.perf_counter() - t0


def run_reference(args, wl, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    The reference is Rust and cannot be compiled in this image (no rustc/cargo), so this arm times
    the oracle port (kind "port") with all host threads on a bounded band of the same workload.
    Under torchrun only rank 0 works; the other ranks exit.
    """
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    arm = CpuArm(wl, args.ref_band_rows)
    for _ in range(args.warmup):
        arm.step(cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step(cores)
    sec = (time.perf_counter() - t0) / args.steps
    rate = arm.n / sec
    sample = f"{arm.desc}, {cores} threads (pixel rows split across threads; every pixel keeps its stream order)"
    out = {
        "impl": "reference",
        "metric": "filtered samples/s into Film",
        "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "filter": wl["filter"], "spp_per_pass": wl["spp"], "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Rust (no rustc here) and has no add_sample: this is oracle/pbrt_oracle.c, the C "
                "restatement of its film path plus pbrt-v3's AddSample, on the host cores",
    }
    emit(out)


# --------------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="exact", choices=["exact", "fma", "atomic"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-band-rows", type=int, default=256, help="band of the film the 1-thread cpu_baseline runs")
    ap.add_argument("--ref-band-rows", type=int, default=1080, help="band of the film the --impl reference arm runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extras", action="store_true", help="(default at N=1) also time the Tier-1 merge / resolve / texture kernels")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-render-c5", action="store_true", help="skip the strong-scaled full render of BASELINE configs[4]")
    ap.add_argument("--render-passes", type=int, default=16, help="16-spp passes of the configs[4] render (16 = 256 spp)")
    args = ap.parse_args()

    # Rank 0 must print exactly one line on stdout.  Libraries (NCCL prints its version on first use)
    # write to fd 1 too, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    # duplicate of the original stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world, emit)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import pbrt_b200 as pb
    from pbrt_b200 import dist as pdist
    from pbrt_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libpbrt_b200 has no CPU fallback")
    # host side of the end-to-end path: keep this rank's threads (and, by first touch, its pinned sample buffers)
    # on the cores next to its GPU; a no-op on single-node hosts or when NVML cannot tell
    bound_cores = pb.bind_host_to_device_numa(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pb.init(local_rank)
    stream = torch.cuda.Stream()          # one explicit stream for the library, the events and NCCL
    torch.cuda.set_stream(stream)
    pb.set_stream(stream.cuda_stream)

    # ---- the film and this rank's shard -------------------------------------------------
    W, H1 = wl["res"]
    H = H1 * world if args.scaling == "weak" else H1
    spp = wl["spp"]
    cls = {"gaussian": pb.GaussianFilter, "mitchell": pb.MitchellFilter, "lanczos": pb.LanczosSincFilter,
           "triangle": pb.TriangleFilter, "box": pb.BoxFilter}[wl["filter"]]
    if wl["filter"] == "mitchell":
        filt = cls(wl["radius"], wl["p0"], wl["p1"])
    elif wl["filter"] in ("gaussian", "lanczos"):
        filt = cls(wl["radius"], wl["p0"])
    else:
        filt = cls(wl["radius"])
    film = pb.Film.new([W, H], [[0, 0], [1, 1]], filt, 35.0, "bench.pfm", 1.0, float("inf"), rank=rank, nranks=world)
    cropped, owned = film.cropped_pixel_bounds, film.owned_pixel_bounds
    full_sb = cropped  # App. C: samples over the cropped pixel bounds
    my_sb = pdist.shard_sample_bounds(full_sb, (owned.p_min.y, owned.p_max.y), wl["radius"][1])
    xy_d, rgbw_d, n_local = synth.samples(my_sb.as4(), spp, seed=1, index_bounds=full_sb.as4())
    n_unique_total = max(cropped.area(), 0) * spp          # every sample counted once
    n_owned = max(owned.area(), 0) * spp
    mode = {"exact": pb.SPLAT_EXACT, "fma": pb.SPLAT_FMA, "atomic": pb.SPLAT_ATOMIC}[args.mode]
    sb_list = [[my_sb.p_min.x, my_sb.p_min.y], [my_sb.p_max.x, my_sb.p_max.y]]

    def step():
        film.add_samples_tile(sb_list, spp, xy_d, rgbw_d, mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    film.check()
    barrier()

    # ---- timed region: exactly K steps, device time, max over ranks ----------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = pb.launch_count()
    # Two events around the K launches and nothing between them: a renderer issues its passes back to back, and an
    # event record between two launches would switch off the programmatic dependent launch that lets pass i+1 fill
    # the SM slots pass i's early CTAs leave (splat_class.cu).  Average launch duration = region / K.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
    ev[1].record(stream)
    barrier()
    clocks = sampler.finish()
    launches = pb.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[1])
    film.check()
    # beside it, outside the timed region: launches bracketed one by one (no overlap between them) — the duration an
    # ncu launch list or a single isolated call sees
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    iso[0].record(stream)
    for i in range(10):
        step()
        iso[i + 1].record(stream)
    barrier()
    per_step = [iso[i].elapsed_time(iso[i + 1]) for i in range(10)]
    # and K back-to-back launches that alternate between two sample buffers, as a renderer that generates new samples for
    # every pass would issue them: a launch that reads other buffers than the one ahead of it waits for that grid before
    # its first load (splat_class.cu), so only its table staging overlaps
    alt_ms = alt_overlap_ms = None
    if world == 1:
        try:
            xy_b, rgbw_b, _ = synth.samples(my_sb.as4(), spp, seed=2, index_bounds=full_sb.as4())
            bufs = [(xy_d, rgbw_d), (xy_b, rgbw_b)]
            for i in range(4):
                film.add_samples_tile(sb_list, spp, *bufs[i & 1], mode)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for i in range(args.steps):
                film.add_samples_tile(sb_list, spp, *bufs[i & 1], mode)
            a1.record(stream)
            barrier()
            alt_ms = a0.elapsed_time(a1) / args.steps
            film.check()
            # the same with pbrt_b200_overlap_passes(1): the caller vouches that a pass's samples were complete before
            # the previous pass was issued (true here: both buffers exist since before the loop)
            pb.overlap_passes(True)
            for i in range(4):
                film.add_samples_tile(sb_list, spp, *bufs[i & 1], mode)
            barrier()
            a0.record(stream)
            for i in range(args.steps):
                film.add_samples_tile(sb_list, spp, *bufs[i & 1], mode)
            a1.record(stream)
            barrier()
            alt_overlap_ms = a0.elapsed_time(a1) / args.steps
            pb.overlap_passes(False)
            film.check()
            del xy_b, rgbw_b, bufs
        except Exception:
            alt_ms = None
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = n_unique_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant (only) kernel of the step, this rank --------------------
    peak, peak_src = measured_peak_gbs()
    kern_ms = total_ms / args.steps  # one launch per step, back to back: region / K = average launch duration
    alg_bytes = n_local * 24 + max(owned.area(), 0) * 32
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": measured_traffic(args.workload, args.mode) if world == 1 else None, "peak_source": peak_src,
        "kernel": "splat_atomic_kernel" if args.mode == "atomic" else ("splat_class_kernel" if wl["radius"][0] in (2.0, 4.0) and spp <= 32 else "splat_window_kernel"),
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms, "kernel_ms_isolated": sum(per_step) / len(per_step), "kernel_ms_alternating_buffers": alt_ms, "kernel_ms_alternating_buffers_overlap_passes": alt_overlap_ms,
        "note": "24 B/sample read + 32 B/film pixel RMW per launch; the splat is bound by shared-memory load latency and instruction issue, not by HBM (DESIGN.md section 5, profiles/)",
    }

    # ---- final assembly (once per render, not per step) ----------------------------------
    # (a) baseline: resolve kernel, then one NCCL all_gather_into_tensor
    pdist.assemble_film_rgb(film, 1.0)  # warm NCCL up
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    frame = pdist.assemble_film_rgb(film, 1.0)
    a1.record(stream)
    barrier()
    assemble_ms = a0.elapsed_time(a1)
    assert tuple(frame.shape) == (H, W, 3)
    # (b) fused: one kernel per rank resolves its rows and stores them into every rank's frame over NVLink
    fx = pdist.FrameExchange(film)
    fx.assemble(1.0)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    fx.launch(1.0)
    b1.record(stream)
    barrier()
    fused = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(fused, op=dist.ReduceOp.MAX)
    assemble_fused_ms = float(fused.item())
    frame2 = fx.assemble(1.0)
    assemble_identical = bool(torch.equal(frame2, frame))
    fx.close()

    # ---- BASELINE configs[4], the whole render, strong-scaled, one timed region ----------------
    render = render3 = None
    if not args.no_render_c5 and args.mode != "atomic":
        try:
            render = render_c5(pb, pdist, synth, torch, dist, stream, rank, world, mode, args.render_passes)
        except Exception as e:  # the headline line must still be printed
            render = {"error": f"{type(e).__name__}: {e}"}
        try:  # BASELINE configs[2] the same way (64 spp = 4 passes)
            render3 = render_c5(pb, pdist, synth, torch, dist, stream, rank, world, mode, 4, config="c3")
        except Exception as e:
            render3 = {"error": f"{type(e).__name__}: {e}"}

    # ---- end to end through the public API with host buffers ------------------------------
    e2e = None
    if not args.no_e2e:
        # AddSample's argument shape: position, radiance and sample weight as separate streams.  The synthetic
        # workload has every sample weight = 1, which the API expresses as "no weight stream":
        # 20 B per sample cross PCIe instead of 24.
        rgbw_h = rgbw_d.to_numpy(np.float32, (n_local, 4))
        weights_all_one = bool((rgbw_h[:, 3] == 1.0).all())
        hxy = pb.PinnedBuffer(np.float32, (n_local, 2))
        hrgb = pb.PinnedBuffer(np.float32, (n_local, 3))
        hsw = None if weights_all_one else pb.PinnedBuffer(np.float32, (n_local,))
        hout = pb.PinnedBuffer(np.float32, (max(owned.area(), 0), 3))
        hxy.array[:] = xy_d.to_numpy(np.float32, (n_local, 2))
        hrgb.array[:] = rgbw_h[:, :3]
        if hsw is not None:
            hsw.array[:] = rgbw_h[:, 3]
        del rgbw_h

        def e2e_step():
            # every step uploads its samples from pinned host memory and reads its resolved frame back; both
            # transfers are enqueued (PBRT_MEM_PINNED_ASYNC) so that step i+1's upload overlaps step i's kernels
            film.add_samples_tile_rgb(sb_list, spp, hxy.array, hrgb.array, None if hsw is None else hsw.array,
                                      mode, pinned_async=True)
            film.resolve_rgb(1.0, out=hout.array, pinned_async=True)

        for _ in range(2):
            e2e_step()
        pb.synchronize()
        k = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        pb.synchronize()   # all uploads, kernels and read-backs of the k steps have completed
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the host link's own ceiling, measured the same way in the same run: every rank copies the same number of
        # bytes from pinned memory with plain cudaMemcpyAsync, all ranks at once (nothing else running)
        h2d_bytes = n_local * (20 if weights_all_one else 24)
        hsrc = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
        ddst = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
        ddst.copy_(hsrc, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(5):
            ddst.copy_(hsrc, non_blocking=True)
        c1.record(stream)
        barrier()
        ct = torch.tensor([c0.elapsed_time(c1) / 5.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        h2d_ceiling = world * h2d_bytes / (float(ct.item()) * 1e-3) / 1e9
        h2d_achieved = world * h2d_bytes / float(dt.item()) / 1e9
        del hsrc, ddst
        e2e = {"value": n_unique_total / float(dt.item()), "unit": "samples/s",
               "h2d_gbs": h2d_achieved, "h2d_ceiling_gbs": h2d_ceiling, "frac_of_h2d_ceiling": h2d_achieved / h2d_ceiling,
               "h2d_ceiling_note": "aggregate over all ranks of plain pinned cudaMemcpyAsync uploads of the step's bytes, every rank at once, measured in this run",
               "h2d_bytes_per_step": n_local * (20 if weights_all_one else 24), "d2h_bytes_per_step": max(owned.area(), 0) * 12,
               "ms_per_step": float(dt.item()) * 1e3,
               "api": "Film.add_samples_tile_rgb(pinned host xy, rgb, sample_weight=%s) + Film.resolve_rgb(pinned host out), transfers enqueued, one synchronize after the K steps" % ("None: all weights are 1" if weights_all_one else "pinned host stream")}
        film.check()

    # ---- secondary kernels (Tier 1: merge / resolve / constant texture) -------------------
    extras = None
    if rank == 0 and not args.no_extras and (args.extras or world == 1):
        try:
            extras = time_extras(pb, synth, film, torch, stream, peak, xy_d, rgbw_d, spp)
        except Exception as e:  # the headline line must still be printed
            extras = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it: rank 0, N = 1 only --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(wl, args.cpu_band_rows)
        arm.step(1)
        secs = [arm.step(1) for _ in range(3)]
        cpu = {"value": arm.n / min(secs), "unit": "samples/s", "cores": 1, "kind": "port",
               "sample": f"{arm.desc}, best of 3, oracle/pbrt_oracle.c on one thread (the reference is single-threaded); "
                         f"host has {os.cpu_count()} cores"}

    if rank == 0:
        out = {
            "metric": "filtered samples/s into Film", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"] + (f", weak-scaled to 1920x{H} over {world} row shards" if world > 1 and args.scaling == "weak" else ""),
                       "film": [W, H], "spp_per_pass": spp, "filter": wl["filter"], "radius": list(wl["radius"]),
                       "mode": args.mode, "samples_per_step": n_unique_total, "samples_per_rank_incl_halo": n_local,
                       "cache": "inputs_exceed_l2" if n_local * 24 > 126e6 else "inputs_fit_l2",
                       "parallelism": f"rows x{world}", "host_cores_bound": (len(bound_cores) if bound_cores else None), "tier": "extension (no reference parity): splat; merge+resolve are Tier 1"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
            "assemble": {"resolve_then_nccl_allgather_ms": assemble_ms, "fused_resolve_peer_store_ms": assemble_fused_ms,
                         "frames_identical": assemble_identical, "bytes_per_rank": max(owned.area(), 0) * 12 * world},
            "percent_of_hbm_peak": 100.0 * achieved / peak,
            "render_c5": render, "render_c3": render3,
        }
        if extras:
            out["extras"] = extras
            # the second half of BASELINE.json's metric ("+ texture lookups/s ... % HBM peak"), N = 1 only
            tf, tr = extras.get("texture_constant_f32"), extras.get("texture_constant_rgb")
            if isinstance(tf, dict) and "lookups_per_s" in tf:
                out["texture_lookups_per_s"] = {"constant_f32": tf["lookups_per_s"], "constant_f32_frac_of_hbm_peak": tf["frac"],
                                                "constant_rgb": tr["lookups_per_s"], "constant_rgb_frac_of_hbm_peak": tr["frac"]}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def render_c5(pb, pdist, synth, torch, dist, stream, rank, world, mode, passes=16, config="c5"):
    """BASELINE configs[4] end to end on the device, STRONG-scaled: the 7680x4320 film row-sharded over `world` GPUs, Lanczos-sinc
    r=4, 256 spp as `passes` pixel-major passes of 16 spp, then the final assembly (fused resolve + NVLink peer stores into
    every rank's full frame) — all inside ONE timed region, device time, max over ranks.  Every pass re-reads the same
    resident 16-spp stream of the shard (12.7 GB at N=1, far above L2): the work per pass is that of a fresh stream, and the
    204 GB of 256 distinct spp would not fit.  At N>1 the block also times one-to-all routing of a pass held by rank 0."""
    # config "c3" = BASELINE configs[2] the same way: 3840x2160, Mitchell r=2, 64 spp as `passes` (4) passes of 16
    if config == "c3":
        W, H, spp, radius = 3840, 2160, 16, 2.0
        filt = pb.MitchellFilter((2.0, 2.0), 1.0 / 3.0, 1.0 / 3.0)
        label = "3840x2160 film, Mitchell r=2, %d spp as %d pixel-major passes of 16 (BASELINE configs[2])"
    else:
        W, H, spp, radius = 7680, 4320, 16, 4.0
        filt = pb.LanczosSincFilter((4.0, 4.0), 3.0)
        label = "7680x4320 film, Lanczos-sinc r=4, %d spp as %d pixel-major passes of 16 (BASELINE configs[4])"
    film = pb.Film.new([W, H], [[0, 0], [1, 1]], filt, 35.0, "render_%s.pfm" % config, 1.0, float("inf"), rank=rank, nranks=world)
    cropped, owned = film.cropped_pixel_bounds, film.owned_pixel_bounds
    sb = pdist.shard_sample_bounds(cropped, (owned.p_min.y, owned.p_max.y), radius)
    xy_d, rgbw_d, n_local = synth.samples(sb.as4(), spp, seed=1, index_bounds=cropped.as4())
    sbl = [[sb.p_min.x, sb.p_min.y], [sb.p_max.x, sb.p_max.y]]
    fx = pdist.FrameExchange(film)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    film.add_samples_tile(sbl, spp, xy_d, rgbw_d, mode)   # warm-up: one pass and one assembly
    fx.assemble(1.0)
    film.clear()
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    launches0 = pb.launch_count()
    e[0].record(stream)
    for _ in range(passes):
        film.add_samples_tile(sbl, spp, xy_d, rgbw_d, mode)
    e[1].record(stream)
    fx.launch(1.0)
    e[2].record(stream)
    barrier()                                              # every peer's stores have landed: all frames complete
    launches = pb.launch_count() - launches0
    film.check()
    t = torch.tensor([e[0].elapsed_time(e[2]), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, splat_ms, assemble_ms = (float(v) for v in t)
    n_total = W * H * spp * passes
    out = {"workload": (label % (spp * passes, passes)) + ", rows sharded over %d GPU(s), then fused resolve + peer-store assembly "
                       "of the full frame on every rank" % world,
           "scaling": "strong", "n_gpus": world, "ms": total_ms, "splat_ms": splat_ms, "assemble_ms": assemble_ms,
           "assemble_share": assemble_ms / total_ms, "samples": n_total, "samples_per_s": n_total / (total_ms * 1e-3),
           "gpu_launches": int(launches), "mode": {pb.SPLAT_EXACT: "exact", pb.SPLAT_FMA: "fma"}.get(mode, str(mode)),
           "samples_per_rank_per_pass_incl_halo": n_local, "timing": "CUDA events on the launching stream, max over ranks",
           "note": "each pass re-reads the same resident 16-spp stream (exceeds L2)"}
    fx.close()
    if world > 1 and config == "c5":
        # one-to-all routing: rank 0 holds a whole 16-spp pass of a 1920-row band and routes it to the owning shards
        from pbrt_b200.dist import _DeviceArray
        band = pb.Bounds2i.raw(0, 0, W, 1080)
        bsb = pb.Bounds2i.raw(-4, -4, W + 4, 1084)
        if rank == 0:
            bxy_d, brgbw_d, bn = synth.samples(bsb.as4(), spp, seed=1)
            bxy = torch.as_tensor(_DeviceArray(bxy_d.ptr, (bn, 2)), device="cuda")
            brgbw = torch.as_tensor(_DeviceArray(brgbw_d.ptr, (bn, 4)), device="cuda")
            rows = (bsb.p_min.y, bsb.p_max.y)
        else:
            bn = 0
            bxy = torch.empty((0, 2), dtype=torch.float32, device="cuda")
            brgbw = torch.empty((0, 4), dtype=torch.float32, device="cuda")
            rows = (bsb.p_max.y, bsb.p_max.y)
        pdist.route_samples(bxy, brgbw, rows, bsb, spp, band, (4.0, 4.0), rank, world)   # warm-up (NCCL channels)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        lxy, lrgbw, lsb = pdist.route_samples(bxy, brgbw, rows, bsb, spp, band, (4.0, 4.0), rank, world)
        r1.record(stream)
        barrier()
        rt = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        n_band = bsb.area() * spp
        out["route_one_to_all"] = {"ms": float(rt.item()), "samples": n_band, "samples_per_s": n_band / (float(rt.item()) * 1e-3),
                                   "bytes_from_rank0": int(n_band * 24 * (world - 1) / world),
                                   "note": "7688x1088-pixel band, 16 spp, all on rank 0; route_samples = NCCL send/recv of row slices, halo rows sent to both neighbours"}
        del lxy, lrgbw
    film.close()
    return out


def time_extras(pb, synth, film, torch, stream, peak, xy_main=None, rgbw_main=None, spp_main=16):
    """Tier-1 kernels (reference-backed): merge_film_tile (48 B/tile px), write_image's resolve (40 B/px),
    ConstantTexture lookups (4 / 12 B).  Run on a 7680x4320 film so that every working set (>= 531 MB)
    exceeds the 126 MB L2 and the numbers are HBM numbers."""
    import numpy as np

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def entry(units, bytes_per_unit, ms, unit_name, **kw):
        gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
        return {unit_name: units / (ms * 1e-3), "GB/s": gbs, "frac": gbs / peak, "ms": ms, "bytes_per_unit": bytes_per_unit, **kw}

    out = {"film": [7680, 4320]}
    big = pb.Film.new([7680, 4320], [[0, 0], [1, 1]], pb.GaussianFilter((2.0, 2.0), 2.0), 35.0, "extras.pfm", 1.0, float("inf"))
    ob = big.owned_pixel_bounds
    npx = ob.area()
    # a8: one whole-frame tile
    buf, offsets, total = synth.tiles([npx], seed=1)
    ms = timed(lambda: big.merge_tile_raw(ob, buf))
    out["merge_film_tile"] = entry(npx, 48, ms, "tile_px_per_s")
    del buf
    # a8: 32x32 sample tiles with their 2-pixel halos, ONE launch, tile order preserved per pixel
    sb = big.get_sample_bounds().as4()
    tbs, counts = [], []
    for y in range(sb[1], sb[3], 32):
        for x in range(sb[0], sb[2], 32):
            tb, cnt = big._tile_bounds([[x, y], [min(x + 32, sb[2]), min(y + 32, sb[3])]])
            tbs.append(tb.as4())
            counts.append(cnt)
    buf2, off2, tot2 = synth.tiles(counts, seed=1)
    b2 = np.asarray(tbs, dtype=np.int32)
    ms = timed(lambda: big.merge_tiles_raw(b2, off2, buf2, tot2), reps=5)
    out["merge_film_tiles_batched"] = entry(tot2, 48, ms, "tile_px_per_s", tiles=len(tbs),
                                            note="36x36-pixel tiles with overlapping halos; per-cell index cached after the first call")
    del buf2
    # a10: resolve
    rgb = torch.empty((npx, 3), dtype=torch.float32, device="cuda")
    ms = timed(lambda: big.resolve_rgb(1.0, out=rgb))
    out["resolve_rgb"] = entry(npx, 40, ms, "px_per_s")
    del rgb
    rgb8 = torch.empty((npx, 3), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: big.resolve_rgb8(1.0, out=rgb8))
    out["resolve_rgb8"] = entry(npx, 31, ms, "px_per_s", note="fused gamma_correct + to_byte (imageio.rs:66-68)")
    del rgb8
    big.close()
    # a16 the way a renderer feeds it: 16x16-sample tiles (pbrt's tile size), each accumulated separately and
    # merged in order — one call, two launches
    try:
        from pbrt_b200.dist import _DeviceArray

        W, H = film.cropped_pixel_bounds.p_max.x, film.cropped_pixel_bounds.p_max.y
        spp = spp_main
        ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
        ty, tx = ys // 16, xs // 16
        tiles_x = (W + 15) // 16
        tile_id = (ty * tiles_x + tx).reshape(-1)
        order = torch.argsort(tile_id, stable=True)          # pixels grouped by tile, row-major inside a tile
        sidx = (order[:, None] * spp + torch.arange(spp, device="cuda")[None, :]).reshape(-1)
        n_s = W * H * spp
        xy_t = torch.as_tensor(_DeviceArray(xy_main.ptr, (n_s, 2)), device="cuda")[sidx].contiguous()
        rgbw_t = torch.as_tensor(_DeviceArray(rgbw_main.ptr, (n_s, 4)), device="cuda")[sidx].contiguous()
        sbs = np.asarray([(x, y, min(x + 16, W), min(y + 16, H)) for y in range(0, H, 16) for x in range(0, W, 16)], dtype=np.int32)
        soffs = np.concatenate([[0], np.cumsum((sbs[:, 2] - sbs[:, 0]).astype(np.int64) * (sbs[:, 3] - sbs[:, 1]) * spp)[:-1]])
        small = pb.Film.new([W, H], [[0, 0], [1, 1]], film.filter, 35.0, "extras_tiles.pfm", 1.0, float("inf"))
        ms = timed(lambda: small.add_samples_tiles(sbs, spp, xy_t, rgbw_t, soffs), reps=5)
        small.check()
        out["splat_tiles_16x16"] = {"samples_per_s": n_s / (ms * 1e-3), "ms": ms, "tiles": len(sbs),
                                    "note": "whole call: host tile descriptors + upload, splat into per-tile buffers, ordered merge"}
        small.close()
        del xy_t, rgbw_t, sidx, order
    except Exception as e:
        out["splat_tiles_16x16"] = {"error": f"{type(e).__name__}: {e}"}
    # the shapes the class kernel does not serve: the reference's own test filter (box r = 8: generic gather) and a
    # pixel-major stream of 64 spp handed over in one call (window kernel on 32-column strips)
    try:
        from pbrt_b200 import synth as _synth
        for key, filt, res, spp_x, note in (
                ("splat_box_r8_generic", pb.BoxFilter([8.0, 8.0]), (1920, 1080), 4,
                 "box r=8 (src/core/film.rs:505-521), 4 spp: splat_gather_generic_kernel, thread per output pixel"),
                ("splat_mitchell_spp64_one_call", pb.MitchellFilter([2.0, 2.0], 1.0 / 3.0, 1.0 / 3.0), (3840, 540), 64,
                 "Mitchell r=2, 64 spp in ONE pixel-major call (a quarter-height band of configs[2]): splat_window_kernel on 32-column strips")):
            f2 = pb.Film.new(list(res), [[0, 0], [1, 1]], filt, 35.0, "extras_x.pfm", 1.0, float("inf"))
            b4 = f2.cropped_pixel_bounds.as4()
            xy_x, rgbw_x, n_x = _synth.samples(b4, spp_x, seed=2)
            sbl = [[b4[0], b4[1]], [b4[2], b4[3]]]
            ms = timed(lambda: f2.add_samples_tile(sbl, spp_x, xy_x, rgbw_x, pb.SPLAT_EXACT), reps=3)
            f2.check()
            out[key] = {"samples_per_s": n_x / (ms * 1e-3), "ms": ms, "film": list(res), "spp": spp_x, "note": note}
            f2.close()
            del xy_x, rgbw_x
    except Exception as e:
        out["splat_other_shapes"] = {"error": f"{type(e).__name__}: {e}"}
    # a14: 1e8 lookups, alternating between two buffers so that no launch finds its lines in L2 (2 x 400 MB, 2 x 1.2 GB)
    n = 100_000_000
    t1 = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
    tex = pb.ConstantTexture(10.0)
    turn = [0]

    def alt(bufs, fn):
        turn[0] ^= 1
        fn(bufs[turn[0]])

    ms = timed(lambda: alt(t1, lambda b: tex.evaluate_batch(n, out=b)))
    out["texture_constant_f32"] = entry(n, 4, ms, "lookups_per_s")
    ms = timed(lambda: alt(t1, lambda b: b.fill_(10.0)))
    out["texture_constant_f32"]["torch_fill_GB/s"] = n * 4 / (ms * 1e-3) / 1e9
    del t1
    t3 = [torch.empty((n, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
    tex3 = pb.ConstantTexture((1.0, 0.0, 0.0))
    ms = timed(lambda: alt(t3, lambda b: tex3.evaluate_batch(n, out=b)))
    out["texture_constant_rgb"] = entry(n, 12, ms, "lookups_per_s")
    ms = timed(lambda: alt(t3, lambda b: b.fill_(1.0)))
    out["texture_constant_rgb"]["torch_fill_GB/s"] = n * 12 / (ms * 1e-3) / 1e9
    out["texture_note"] = "two output buffers alternated per launch (working set 2 x 400 MB / 2 x 1.2 GB >> 126 MB L2); torch_fill = eager torch.Tensor.fill_ on the same buffers"
    return out


if __name__ == "__main__":
    main()
