"""Row-sharded films on the GPU(s): every shard bit-identical to the matching rows of the single film.
Runs on one GPU (the shards are independent films); with >1 GPU bench.py --gpus N adds the NCCL all-gather."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_splat_equals_single_film(gpu, orc, world):
    from pbrt_b200 import dist as pdist
    from pbrt_b200 import synth

    res, spp = (200, 120), 4
    filt = gpu.MitchellFilter((2.0, 2.0), 1 / 3, 1 / 3)
    whole = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    full_sb = whole.cropped_pixel_bounds
    xy, rgbw, n = synth.samples(full_sb.as4(), spp, seed=1)
    whole.add_samples_tile(full_sb.as4(), spp, xy, rgbw, gpu.SPLAT_EXACT)
    ref = whole.resolve_rgb(1.0).reshape(res[1], res[0], 3)
    parts = []
    for rank in range(world):
        f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"), rank=rank, nranks=world)
        ob = f.owned_pixel_bounds
        assert (ob.p_min.y, ob.p_max.y) == pdist.shard_rows(full_sb, rank, world)
        sb = pdist.shard_sample_bounds(full_sb, (ob.p_min.y, ob.p_max.y), 2.0)
        sxy, srgbw, sn = synth.samples(sb.as4(), spp, seed=1, index_bounds=full_sb.as4())
        f.add_samples_tile(sb.as4(), spp, sxy, srgbw, gpu.SPLAT_EXACT)
        f.check()
        parts.append(f.resolve_rgb(1.0).reshape(-1, res[0], 3))
    got = np.concatenate(parts, axis=0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
