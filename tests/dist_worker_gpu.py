"""Worker of tests/test_gpu_multi.py: one process per GPU under torch.distributed.run (NCCL).

Rank 0 holds the WHOLE pixel-major stream; pbrt_b200.dist.route_samples delivers every shard its rows plus halo over
NVLink; each rank splats its shard; the frame is assembled twice — resolve + NCCL all-gather (assemble_film_rgb) and the
fused resolve-with-peer-stores kernel (FrameExchange) — and both must equal, bit for bit, the frame of a single
unsharded film that rank 0 renders from the same stream."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import pbrt_b200 as pb
    from pbrt_b200 import dist as pdist
    from pbrt_b200 import synth
    from pbrt_b200.dist import _DeviceArray

    out_path, W, H, spp, fname = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pb.init(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    pb.set_stream(stream.cuda_stream)
    filt = {"gaussian": lambda: pb.GaussianFilter((2.0, 2.0), 2.0), "lanczos": lambda: pb.LanczosSincFilter((4.0, 4.0), 3.0)}[fname]()
    radius = filt.radius()
    film = pb.Film.new([W, H], [[0, 0], [1, 1]], filt, 35.0, "multi.pfm", 1.0, float("inf"), rank=rank, nranks=world)
    cropped = film.cropped_pixel_bounds
    sb = pb.Bounds2i.of(film.get_sample_bounds())
    n_all = sb.area() * spp
    result = {"rank": rank}
    if rank == 0:
        xy_d, rgbw_d, n = synth.samples(sb.as4(), spp, seed=3)
        xy = torch.as_tensor(_DeviceArray(xy_d.ptr, (n, 2)), device="cuda")
        rgbw = torch.as_tensor(_DeviceArray(rgbw_d.ptr, (n, 4)), device="cuda")
        rows = (sb.p_min.y, sb.p_max.y)
    else:
        xy = torch.empty((0, 2), dtype=torch.float32, device="cuda")
        rgbw = torch.empty((0, 4), dtype=torch.float32, device="cuda")
        rows = (sb.p_max.y, sb.p_max.y)
    torch.cuda.synchronize()
    lxy, lrgbw, lsb = pdist.route_samples(xy, rgbw, rows, sb, spp, cropped, radius, rank, world)
    torch.cuda.synchronize()
    ob = film.owned_pixel_bounds
    want_sb = pdist.shard_sample_bounds(sb, (ob.p_min.y, ob.p_max.y), radius[1])
    assert lsb.as4() == want_sb.as4(), (lsb.as4(), want_sb.as4())
    film.add_samples_tile(lsb.as4(), spp, lxy, lrgbw, pb.SPLAT_EXACT)
    film.check()
    frame_nccl = pdist.assemble_film_rgb(film, 1.0)
    fx = pdist.FrameExchange(film)
    frame_fused = fx.assemble(1.0)
    result["frames_identical"] = bool(torch.equal(frame_nccl, frame_fused))
    result["shape"] = list(frame_nccl.shape)
    if rank == 0:
        whole = pb.Film.new([W, H], [[0, 0], [1, 1]], filt, 35.0, "single.pfm", 1.0, float("inf"))
        whole.add_samples_tile(sb.as4(), spp, xy_d, rgbw_d, pb.SPLAT_EXACT)
        whole.check()
        ref = torch.from_numpy(whole.resolve_rgb(1.0).reshape(H, W, 3)).cuda()
        result["equals_single_film"] = bool(torch.equal(frame_nccl.view(torch.int32), ref.view(torch.int32)))
        result["nonzero"] = bool((ref != 0).any())
    # every rank's copy of the frame is the same
    mine = frame_fused.contiguous().view(torch.int32).sum(dtype=torch.int64).reshape(1)
    sums = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(sums, mine)
    result["all_ranks_same_frame"] = bool(all(int(s) == int(sums[0]) for s in sums))
    torch.cuda.synchronize()
    dist.barrier()
    fx.close()
    Path(f"{out_path}.{rank}").write_text(json.dumps(result))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
