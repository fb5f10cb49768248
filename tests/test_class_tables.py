"""Host side of the phase-class splat kernel (pbrt_b200/csrc/splat_class.cu), checked on the CPU box.

The kernel replaces the per-tap index arithmetic of add_sample by a classification of the sample's sub-pixel phase and a
block of precombined weights per class pair.  The LUTs and blocks are built by host code; here they are read back through
a test hook, the kernel's classification is replayed in numpy float32, and the weights it would use for every
(column, row) of the sample's footprint are compared with what the CPU oracle's add_sample deposits for that very
sample — for interval classes, point classes, phases one ulp either side of every class boundary, and pixels next to a
power-of-two coordinate where floor(pd + r) rounds."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import OracleFilm

CF_UP, CF_DOWN, CF_LEFT, CF_RIGHT = 1, 2, 4, 8
FAST_ENTRIES, CARE_ENTRIES = 32, 16
LUT_Y, LUT_X, CARE_Y, CARE_X, Q_OFFSET = 0, 512, 1024, 1536, 2048
SLOW_X, SLOW_Y = 0x40000000, 0x80000000


def class_tables(pb, table, radius):
    from pbrt_b200 import _lib

    geom = (C.c_int32 * 9)()
    t = np.ascontiguousarray(table, dtype=np.float32)
    n = _lib.lib.pbrt_b200_debug_class_tables(t.ctypes.data_as(C.POINTER(C.c_float)), radius, radius, None, 0, geom)
    if n == 0:
        return None, None
    blob = np.zeros(n, dtype=np.uint8)
    _lib.lib.pbrt_b200_debug_class_tables(t.ctypes.data_as(C.POINTER(C.c_float)), radius, radius,
                                          blob.ctypes.data_as(C.POINTER(C.c_uint8)), n, geom)
    names = ["H", "K", "NX", "NXP", "COLB", "BLK", "ROWP", "EOFF", "BYTES"]
    return blob, dict(zip(names, list(geom)))


class Replay:
    """The pre-pass classification and the gather's weight fetch of splat_class_kernel, in numpy float32."""

    def __init__(self, blob, g):
        self.g = g
        self.f32 = blob.view(np.float32)
        self.u32 = blob.view(np.uint32)

    def classify(self, w, axis, careful):
        """-> offset word (>= 0x40000000: no class) exactly as the kernel's pre-pass computes it"""
        K = self.g["K"]
        if not careful:   # fast LUT: cell = floor(K * (w + 0.5)); {point, offset(interval), offset(point)}
            t = np.float32(np.float64(w) * K + 0.5 * K)          # fma: one rounding
            t = min(abs(t), np.float32(FAST_ENTRIES - 1)) if t == t else np.float32(FAST_ENTRIES - 1)
            base = (LUT_X if axis == "x" else LUT_Y) // 4 + int(np.floor(t)) * 4
            return int(self.u32[base + 2]) if w == self.f32[base] else int(self.u32[base + 1])
        t = np.float32(np.float64(w) * K + 0.5 * K + 0.5)
        t = min(abs(t), np.float32(CARE_ENTRIES - 1)) if t == t else np.float32(CARE_ENTRIES - 1)
        base = (CARE_X if axis == "x" else CARE_Y) // 4 + int(np.floor(t)) * 8
        glo, ghi, point = self.f32[base], self.f32[base + 1], self.f32[base + 2]
        ob, op, oa = (int(v) for v in self.u32[base + 4: base + 7])
        if w < glo:
            return ob
        if w > ghi:
            return oa
        return op if w == point else (SLOW_X if axis == "x" else SLOW_Y)

    def weights(self, wx, wy, n, radius):
        """-> None (slow path) or a (ROWS, ROWS) array [row j][column v] of the weights the gather would add."""
        g = self.g
        H, ROWS, LIVE = g["H"], 2 * g["H"] + 1, 2 * g["H"]
        careful = abs(n[0]) < 2.5 or abs(n[1]) < 2.5
        offx, offy = self.classify(wx, "x", careful), self.classify(wy, "y", careful)
        pdx, pdy = np.float32(n[0]) + wx, np.float32(n[1]) + wy
        r = np.float32(radius)
        ok = offx + offy < SLOW_X
        ok = ok and not (wx < 0 and np.float32(n[0]) + r <= np.float32(pdx + r))
        ok = ok and not (wy < 0 and np.float32(n[1]) + r <= np.float32(pdy + r))
        if not ok:
            return None
        s = offx + offy
        fl, meta = s & 15, s & ~15
        out = np.zeros((ROWS, ROWS), dtype=np.float32)
        for v in range(ROWS):  # v = dx + H; the thread of column x = n + dx visits with d = -dx
            if v == 0 and not fl & CF_LEFT:
                continue
            if v == ROWS - 1 and not fl & CF_RIGHT:
                continue
            col = self.f32[(meta + v * g["COLB"]) // 4:][:LIVE]
            assert fl & (CF_UP | CF_DOWN)
            both = (fl & 3) == 3  # phase 0: either kind, plus the row the other kind would add
            cx = (meta - (Q_OFFSET + g["K"] * g["ROWP"])) // g["BLK"]
            if fl & CF_UP:
                out[:LIVE, v] = col
                if both:
                    out[ROWS - 1, v] = self.f32[g["EOFF"] // 4 + cx * ROWS + v]
            if fl & CF_DOWN:
                down = np.zeros(ROWS, dtype=np.float32)
                for i in range(LIVE):
                    down[i + 1] = col[LIVE - 1 - i]
                if both:
                    down[0] = self.f32[g["EOFF"] // 4 + cx * ROWS + v]
                    assert np.array_equal(down.view(np.uint32), out[:, v].view(np.uint32))  # the two readings agree
                out[:, v] = down
        return out


def oracle_weights(orc, of, n, p, H):
    """What add_sample deposits for one sample of unit radiance at p: (ROWS, ROWS) [row][column] around pixel n, and the
    mask of those pixels that lie on the film (the tile is clipped to it)."""
    ROWS = 2 * H + 1
    sb = (n[0], n[1], n[0] + 1, n[1] + 1)
    x0, y0, x1, y1 = of.tile_bounds(sb)
    tile = of.get_film_tile(sb)
    xy = np.array([p], dtype=np.float32)
    rgbw = np.array([[1, 1, 1, 1]], dtype=np.float32)
    orc.orc_ext_tile_add_samples(tile, 1, oracle.fp(xy), oracle.fp(rgbw))
    px = of.tile_pixels(tile).copy()
    assert np.array_equal(px[:, 0], px[:, 3])
    out = np.zeros((ROWS, ROWS), dtype=np.float32)
    on_film = np.zeros((ROWS, ROWS), dtype=bool)
    sub = px[:, 3].reshape(y1 - y0, x1 - x0)
    out[y0 - (n[1] - H): y1 - (n[1] - H), x0 - (n[0] - H): x1 - (n[0] - H)] = sub
    on_film[y0 - (n[1] - H): y1 - (n[1] - H), x0 - (n[0] - H): x1 - (n[0] - H)] = True
    return out, on_film


def phases(K):
    """Phases worth testing on one axis: every point, every midpoint, and the floats around each class boundary."""
    out = []
    for m in range(K + 1):
        p = np.float32(m / K - 0.5)
        out.append(p)
        for step in (1, 2, 3):
            a = p
            b = p
            for _ in range(step):
                a = np.nextafter(a, np.float32(-1))
                b = np.nextafter(b, np.float32(1))
            out += [a, b]
        # the grid a coordinate near 1000 / 4000 quantises to
        for q in (2.0 ** -14, 2.0 ** -12):
            out += [np.float32(p - q), np.float32(p + q)]
    for k in range(K):
        out.append(np.float32((k + 0.5) / K - 0.5))
    rng = np.random.default_rng(7)
    out += list(rng.uniform(-0.5, 0.5, 12).astype(np.float32))
    return [w for w in out if -0.5 <= w <= 0.5]


@pytest.mark.parametrize("name,radius", [("gaussian", 2.0), ("mitchell", 2.0), ("lanczos", 4.0), ("distinct", 2.0),
                                         ("distinct", 4.0)])
def test_class_blocks_equal_oracle_add_sample(pb, orc, name, radius):
    if name == "distinct":  # every entry different: an index mix-up cannot hide behind equal weights
        table = (1.0 + np.arange(256) / 1024.0).astype(np.float32)
    else:
        kind, rad, p0, p1 = oracle.FILTERS[name]
        table = oracle.filter_table(orc, kind, (radius, radius), p0, p1)
    blob, g = class_tables(pb, table, radius)
    assert blob is not None and g["H"] == int(radius) and g["K"] == 16 // int(radius)
    rp = Replay(blob, g)
    H = g["H"]
    res = (1100, 80)
    of = OracleFilm(orc, res, [0, 0, 1, 1], (radius, radius), table)
    ws = phases(g["K"])
    slow = total = 0
    for n in ((37, 21), (1022 - H + 1, 30), (1023, 63 - H), (511, 31), (1, 2), (0, 0), (2, 5), (4, 1)):
        for wx in ws:
            for wy in (ws if n == (37, 21) else ws[:: 7]):
                # the sample whose phases are (wx, wy): p = n + w + 0.5, rounded as a renderer's float would be
                p = (np.float32(np.float32(n[0]) + wx + np.float32(0.5)), np.float32(np.float32(n[1]) + wy + np.float32(0.5)))
                pwx = np.float32(np.float32(p[0] - np.float32(0.5)) - np.float32(n[0]))
                pwy = np.float32(np.float32(p[1] - np.float32(0.5)) - np.float32(n[1]))
                if not (abs(pwx) <= 0.5 and abs(pwy) <= 0.5):
                    continue
                total += 1
                got = rp.weights(pwx, pwy, n, radius)
                if got is None:
                    # never at an ordinary pixel; near the origin only the floats a few ulps off a class point; at a
                    # power-of-two coordinate the samples whose floor(pd + r) rounds up
                    assert n != (37, 21) and n != (511, 31), (n, float(pwx), float(pwy))
                    slow += 1
                    continue
                want, on_film = oracle_weights(orc, of, n, p, H)
                assert np.array_equal(got.view(np.uint32)[on_film], want.view(np.uint32)[on_film]), (n, float(pwx), float(pwy))
    assert total > 2000 and slow < 0.05 * total, (slow, total)


def test_classes_cover_random_phases(pb, orc):
    """A random in-pixel phase is almost never left to the slow path (only floats within a few ulps of a boundary are)."""
    kind, rad, p0, p1 = oracle.FILTERS["gaussian"]
    blob, g = class_tables(pb, oracle.filter_table(orc, kind, rad, p0, p1), 2.0)
    rp = Replay(blob, g)
    rng = np.random.default_rng(3)
    w = rng.uniform(-0.5, 0.5, 20000).astype(np.float32)
    # positions on the 2^-14 grid of a coordinate near 1000 land on class boundaries all the time: still classes
    w[::2] = np.round(w[::2] * 16384) / 16384
    miss = sum(1 for v in w[::2] if rp.classify(v, "x", False) >= SLOW_X)
    miss += sum(1 for v in w if rp.classify(v, "y", True) >= SLOW_X)
    assert miss == 0


@pytest.mark.parametrize("radius", [0.5, 1.0, 1.5, 3.0, 8.0])
def test_other_radii_have_no_class_tables(pb, radius):
    blob, g = class_tables(pb, np.ones(256, dtype=np.float32), radius)
    assert blob is None


# ---- row segments of the one-wave grid (splat_class.cu: class_segment_rows) ----

def segments(y0, rows, cols, segs, per_sm=4, h=2, nsm=148):
    from pbrt_b200 import _lib

    out = (C.c_int32 * (segs + 1))()
    n = _lib.lib.pbrt_b200_debug_class_segments(y0, rows, cols, segs, per_sm, h, nsm, out)
    assert n == segs + 1
    return list(out)


@pytest.mark.parametrize("y0,rows,cols,segs,h", [(0, 1080, 15, 39, 2), (0, 2160, 30, 19, 2), (0, 4320, 60, 9, 4),
                                                  (-3, 67, 1, 8, 2), (1080, 540, 60, 9, 4), (5, 33, 2, 33, 2), (0, 1, 1, 1, 2)])
def test_row_segments_cover_the_rows_once_and_shrink_with_residency_rank(y0, rows, cols, segs, h):
    s = segments(y0, rows, cols, segs, h=h)
    assert s[0] == y0 and s[-1] == y0 + rows
    sizes = [b - a for a, b in zip(s, s[1:])]
    assert min(sizes) >= 1 and sum(sizes) == rows
    # CTAs are placed round-robin by linear block id: a segment's rank on its SM is (segment * cols + x) / SMs
    ranks = [min(3, (i * cols + cols // 2) // 148) for i in range(segs)]
    by_rank = {}
    for r, n in zip(ranks, sizes):
        by_rank.setdefault(r, []).append(n)
    means = [sum(v) / len(v) for _, v in sorted(by_rank.items())]
    if rows >= 8 * segs:  # (tiny grids are all rounding)
        assert all(a >= b for a, b in zip(means, means[1:])), means        # an older CTA runs faster and gets more rows
        assert max(sizes) - min(sizes) <= max(2, round(0.2 * rows / segs)), sizes


def test_row_segments_of_a_grid_smaller_than_the_machine_are_equal():
    sizes = [b - a for a, b in zip(*(lambda s: (s, s[1:]))(segments(0, 640, 4, 16)))]   # 64 CTAs on 148 SMs: all rank 0
    assert max(sizes) - min(sizes) <= 1
