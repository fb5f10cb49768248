"""Constant texture, WEIGHT_LUT, tile generator on the GPU against the oracle."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1024, 100003, 1 << 20])
def test_constant_texture_batches(gpu, orc, kats, n):
    k = kats["constant_texture"]
    got = gpu.ConstantTexture.new(k["float_value"]).evaluate_batch(n)
    want = np.zeros(n, dtype=np.float32)
    orc.orc_constant_texture_eval_f32(1, k["float_value"], n, oracle.fp(want))
    assert got.shape == (n,) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    got3 = gpu.ConstantTexture.new(k["spectrum_value"]).evaluate_batch(n)
    want3 = np.zeros((n, 3), dtype=np.float32)
    orc.orc_constant_texture_eval_rgb(1, oracle.farr(k["spectrum_value"]), n, oracle.fp(want3))
    assert np.array_equal(got3.view(np.uint32), want3.view(np.uint32))
    d = gpu.create_constant_spectrum_texture().evaluate_batch(n)
    assert (d == 1.0).all()


def test_constant_texture_device_output_unaligned(gpu):
    buf = gpu.DeviceBuffer(4 * 4099 * 3 + 64)
    buf.zero()

    class Off:  # a device pointer 4 bytes past a 16-byte boundary
        ptr, nbytes = buf.ptr + 4, 4 * 4099 * 3
    import ctypes as C
    from pbrt_b200 import _lib

    _lib.check(_lib.lib.pbrt_texture_constant_eval_rgb(_lib.f32arr([1.0, 2.0, 3.0]), 4099, C.c_void_p(Off.ptr), 1))
    got = buf.to_numpy(np.float32, (4099 * 3 + 16,))
    assert got[0] == 0 and np.array_equal(got[1:1 + 4099 * 3].reshape(-1, 3), np.tile([1, 2, 3], (4099, 1)).astype(np.float32))
    assert (got[1 + 4099 * 3:] == 0).all()
    _lib.check(_lib.lib.pbrt_texture_constant_eval_f32(7.0, 4099, C.c_void_p(Off.ptr), 1))
    got = buf.to_numpy(np.float32, (4099 + 8,))
    assert got[0] == 0 and (got[1:4100] == 7.0).all() and got[4100] != 7.0


def test_weight_lut_vs_oracle(gpu, orc):
    want = np.zeros(128, dtype=np.float32)
    orc.orc_weight_lut(oracle.fp(want))
    got = gpu.weight_lut()
    # device expf vs glibc expf: within 2 ulp of the larger term; bit-exact is not promised
    assert np.abs(got - want).max() <= 2.4e-7
    assert got[127] == 0.0


def test_tile_generator_vs_oracle(gpu, orc):
    from pbrt_b200 import synth
    from oracle import OracleFilm

    of = OracleFilm(orc, (64, 64), [0, 0, 1, 1], (2, 2), np.ones(256, np.float32))
    sbs = [(0, 0, 16, 16), (16, 0, 32, 16), (48, 48, 64, 64)]
    tiles = [of.get_film_tile(sb) for sb in sbs]
    counts = [orc.orc_tile_pixel_count(t) for t in tiles]
    buf, offsets, total = synth.tiles(counts, seed=1)
    got = buf.to_numpy(np.float32, (total, 4))
    for i, t in enumerate(tiles):
        orc.orc_ext_synth_tile_fill(t, 1, i)
        want = of.tile_pixels(t)
        assert np.array_equal(got[offsets[i]:offsets[i] + counts[i]].view(np.uint32), want.view(np.uint32))
        orc.orc_tile_free(t)


def test_launch_counter_and_device_info(gpu):
    info = gpu.device_info()
    assert info["cc"][0] == 10 and info["sm_count"] >= 100
    a = gpu.launch_count()
    gpu.ConstantTexture(1.0).evaluate_batch(16)
    assert gpu.launch_count() == a + 1
