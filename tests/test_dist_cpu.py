"""Row-sharded path on CPU: 2 processes over gloo, each computing its shard with the oracle, assembled by
pbrt_b200.dist.allgather_rows — the host-side logic of the N>1 path (partition, halo routing, assembly).
The GPU arm of the same logic is exercised by bench.py --gpus N and tests/test_gpu_dist.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, res, spp, name, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    from oracle import OracleFilm
    from pbrt_b200 import dist as pdist
    from pbrt_b200.geometry import Bounds2i

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.load()
    kind, rad, p0, p1 = oracle.FILTERS[name]
    table = oracle.filter_table(o, kind, rad, p0, p1)
    W, H = res
    cropped = Bounds2i.raw(0, 0, W, H)
    y0, y1 = pdist.shard_rows(cropped, rank, world)
    # the shard is a Film whose crop window is its row block (src/core/film.rs:92-101)
    film = OracleFilm(o, res, [0.0, y0 / H, 1.0, y1 / H], rad, table)
    assert film.cropped() == (0, y0, W, y1)
    sb = pdist.shard_sample_bounds(cropped, (y0, y1), rad[1])
    # this rank's slice of the global stream (pixel indices taken over the whole film)
    xy_all, rgbw_all = oracle.synth_samples(o, cropped.as4(), spp, 1)
    lo, hi = (sb.p_min.y * W) * spp, (sb.p_max.y * W) * spp
    film.add_samples_pass(sb.as4(), spp, xy_all[lo:hi], rgbw_all[lo:hi])
    local = torch.from_numpy(film.write_image_rgb(1.0).reshape(y1 - y0, W, 3))
    full = pdist.allgather_rows(local, cropped, rank, world)
    assert tuple(full.shape) == (H, W, 3)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("res,world", [((48, 37), 2), ((32, 30), 3)])
def test_row_sharded_assembly_matches_single_film(orc, tmp_path, res, world):
    import torch.multiprocessing as mp

    import oracle
    from oracle import OracleFilm

    spp, name = 4, "mitchell"
    port = _free_port()
    mp.spawn(_worker, args=(world, port, res, spp, name, str(tmp_path)), nprocs=world, join=True)
    kind, rad, p0, p1 = oracle.FILTERS[name]
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    whole = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    xy, rgbw = oracle.synth_samples(orc, (0, 0, *res), spp, 1)
    whole.add_samples_pass((0, 0, *res), spp, xy, rgbw)
    want = whole.write_image_rgb(1.0).reshape(res[1], res[0], 3)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        # sharding changes nothing: every pixel sees the same samples in the same order
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"rank {r}"


def test_shard_helpers(pb):
    from pbrt_b200 import dist as pdist

    c = pb.Bounds2i.raw(0, 10, 100, 110)
    rows = pdist.all_rows(c, 8)
    assert rows[0][0] == 10 and rows[-1][1] == 110
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert pdist.halo_rows(2.0) == 2 and pdist.halo_rows(0.5) == 1 and pdist.halo_rows(4.0) == 4 and pdist.halo_rows(1.4) == 1
    sb = pdist.shard_sample_bounds(c, rows[3], 2.0)
    assert sb.p_min.y == rows[3][0] - 2 and sb.p_max.y == rows[3][1] + 2
    assert pdist.shard_sample_bounds(c, rows[0], 2.0).p_min.y == 10


# ------------------------------------------------------------------ sample routing

def _route_worker(rank, world, port, res, spp, blocks, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    from pbrt_b200 import dist as pdist
    from pbrt_b200.geometry import Bounds2i

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle.load()
    W, H = res
    cropped = Bounds2i.raw(0, 0, W, H)
    sb = Bounds2i.raw(-2, -2, W + 2, H + 2)          # get_sample_bounds of a radius-2 filter
    xy_all, rgbw_all = oracle.synth_samples(o, sb.as4(), spp, 1)
    per_row = (W + 4) * spp
    a, b = blocks[rank]                               # the rows this rank happens to hold
    lo, hi = (a - sb.p_min.y) * per_row, (b - sb.p_min.y) * per_row
    xy, rgbw, got_sb = pdist.route_samples(torch.from_numpy(xy_all[lo:hi].copy()), torch.from_numpy(rgbw_all[lo:hi].copy()),
                                           (a, b), sb, spp, cropped, (2.0, 2.0), rank, world)
    y0, y1 = pdist.shard_rows(cropped, rank, world)
    want_sb = pdist.shard_sample_bounds(sb, (y0, y1), 2.0)
    assert got_sb.as4() == want_sb.as4(), (got_sb.as4(), want_sb.as4())
    wlo, whi = (want_sb.p_min.y - sb.p_min.y) * per_row, (want_sb.p_max.y - sb.p_min.y) * per_row
    assert np.array_equal(xy.numpy().view(np.uint32), xy_all[wlo:whi].view(np.uint32))
    assert np.array_equal(rgbw.numpy().view(np.uint32), rgbw_all[wlo:whi].view(np.uint32))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,blocks", [
    (2, [(-2, 42), (42, 42)]),              # rank 0 holds the whole stream
    (3, [(-2, 9), (9, 10), (10, 42)]),      # uneven blocks that do not line up with the shards
    (3, [(30, 42), (-2, 5), (5, 30)]),      # blocks in a different order than the shards
])
def test_route_samples_delivers_each_shard_its_rows_and_halo(orc, tmp_path, world, blocks):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_route_worker, args=(world, port, (24, 40), 2, blocks, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_route_plan_matches_shard_sample_bounds(pb):
    from pbrt_b200 import dist as pdist

    c = pb.Bounds2i.raw(0, 10, 100, 110)
    sb = pb.Bounds2i.raw(-4, 6, 104, 114)
    for n in (1, 2, 3, 8):
        plan = pdist.route_plan(sb, c, (4.0, 4.0), n, (sb.p_min.y, sb.p_max.y))
        for g in range(n):
            want = pdist.shard_sample_bounds(sb, pdist.shard_rows(c, g, n), 4.0)
            assert plan[g] == (want.p_min.y, want.p_max.y)
    # a source holding rows [20, 30) owes a far shard nothing
    assert all(b == a for a, b in pdist.route_plan(sb, c, (4.0, 4.0), 8, (20, 30))[4:])
