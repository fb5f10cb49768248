"""Host-side mirror: geometry semantics and the image containers (no GPU)."""
import numpy as np
import pytest


def test_bounds2i_semantics(pb, kats):
    k = kats["bounds2i"]
    B = pb.Bounds2i
    b = B.of([[2, 2], [4, 4]])
    assert b.inside_exclusive((2, 2)) and not b.inside_exclusive((4, 4))
    assert B.intersect(B.of([[1, 1], [3, 3]]), B.of([[2, 2], [4, 4]])) == B.of([[2, 2], [3, 3]])
    inv = B.intersect(B.of([[1, 1], [2, 2]]), B.of([[3, 3], [4, 4]]))
    assert inv.as4() == tuple(k["intersect_disjoint"]["result"])
    assert inv.area() == 1 and list(inv.iter()) == []
    assert [tuple(p) for p in b.iter()] == [tuple(p) for p in k["iter"]["points"]]
    assert B.of([[5, 4], [3, 2]]).as4() == (3, 2, 5, 4)


def _gradient():
    res = 64
    ys, xs = np.mgrid[0:res, 0:res]
    px = np.stack([xs / np.float32(res), ys / np.float32(res), np.ones_like(xs, dtype=np.float32)], axis=-1)
    return px.astype(np.float32).reshape(-1, 3), res


def test_roundtrip_pfm(pb, tmp_path):
    """src/core/imageio.rs:362-390 — bit-exact."""
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "g.pfm")
    imageio.write_image(name, px.reshape(-1), [[0, 0], [res, res]], (res, res))
    got, r = imageio.read_image(name)
    assert (r.x, r.y) == (res, res)
    assert np.array_equal(got.view(np.uint32), px.view(np.uint32))


def test_roundtrip_png(pb, orc, tmp_path):
    """src/core/imageio.rs:325-360 — compare after to_byte / 255."""
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "g.png")
    imageio.write_image(name, px.reshape(-1), [[0, 0], [res, res]], (res, res))
    got, r = imageio.read_image(name)
    want = np.array([orc.orc_to_byte(float(v)) for v in px.reshape(-1)], dtype=np.float32) / np.float32(255)
    assert (r.x, r.y) == (res, res)
    assert np.array_equal(got.reshape(-1), want)


def test_pfm_bytes_match_oracle(pb, orc, tmp_path):
    import ctypes as C

    import oracle
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "o.pfm")
    imageio.write_pfm(name, px, (res, res))
    need = orc.orc_pfm_encode(oracle.fp(px), res, res, None, 0)
    buf = (C.c_uint8 * need)()
    orc.orc_pfm_encode(oracle.fp(px), res, res, buf, need)
    assert open(name, "rb").read() == bytes(buf)


def test_host_to_byte_matches_oracle(pb, orc):
    from pbrt_b200 import imageio

    v = np.linspace(-0.1, 1.2, 4001, dtype=np.float32)
    got = imageio.to_byte(v)
    want = np.array([orc.orc_to_byte(float(x)) for x in v], dtype=np.uint8)
    # numpy's powf and glibc's differ in the last ulp at most: at most a stray LSB on a tie
    assert (np.abs(got.astype(int) - want.astype(int)) <= 1).all()
    assert (got != want).mean() < 1e-3


def test_unknown_extension(pb, tmp_path):
    from pbrt_b200 import imageio

    with pytest.raises(ValueError):
        imageio.write_image(str(tmp_path / "a.xyz"), np.zeros(3, np.float32), [[0, 0], [1, 1]])
    with pytest.raises(NotImplementedError):
        imageio.read_image(str(tmp_path / "a.exr"))


def test_constant_texture_scalar_evaluate_is_host_side(pb, kats):
    k = kats["constant_texture"]
    si = pb.SurfaceInteraction()
    assert pb.ConstantTexture.new(k["float_value"]).evaluate(si) == 10.0
    assert list(pb.ConstantTexture.new(k["spectrum_value"]).evaluate(si)) == k["spectrum_value"]
    assert pb.create_constant_float_texture().evaluate(si) == k["float_default"]
    assert list(pb.create_constant_spectrum_texture().evaluate(si)) == k["spectrum_default"]
    assert pb.create_constant_float_texture(None, {"value": 10.0}).evaluate(si) == 10.0
    assert "ConstantTexture{" in repr(pb.ConstantTexture(10.0))


def test_to_byte_threshold_table_agrees_with_glibc(orc):
    """The device's to_byte is driven by a threshold table generated without libm (correctly rounded powf).
    Check it against the oracle (glibc powf, what Rust's f32::powf calls on Linux) at every threshold and
    at the float just below it."""
    import re
    import struct
    from pathlib import Path

    txt = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "to_byte_table.inc").read_text()
    thr = [int(h, 16) for h in re.findall(r"0x([0-9a-f]{8})u", txt)]
    assert len(thr) == 256 and thr[0] == 0 and thr == sorted(thr)
    for k in range(1, 256):
        v = struct.unpack("<f", struct.pack("<I", thr[k]))[0]
        below = struct.unpack("<f", struct.pack("<I", thr[k] - 1))[0]
        assert orc.orc_to_byte(v) == k, (k, v)
        assert orc.orc_to_byte(below) == k - 1, (k, below)


def test_to_byte_slices_hold_at_most_one_threshold():
    """The resolve kernel looks to_byte up by the top bits of the float (exponent + 6 mantissa bits, a "slice")
    and settles the result with ONE comparison against the next threshold.  That is exact only if no slice
    contains two thresholds and the range below the first slice / above the last maps to 0 / 255."""
    import re
    from pathlib import Path

    txt = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "to_byte_table.inc").read_text()
    thr = [int(h, 16) for h in re.findall(r"0x([0-9a-f]{8})u", txt)][1:]
    cu = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "film.cu").read_text()
    shift = int(re.search(r"SLICE_SHIFT = (\d+);", cu).group(1))
    first = int(re.search(r"SLICE_FIRST_BITS = 0x([0-9a-f]+)u;", cu).group(1), 16)
    one = 0x3F800000
    assert first < thr[0], "inputs below the first slice must all map to byte 0"
    assert thr[-1] <= one, "1.0 must already map to 255"
    for start in range(first, one, 1 << shift):
        inside = [t for t in thr if start < t < start + (1 << shift)]
        assert len(inside) <= 1, (hex(start), inside)
