"""A second, independent restatement of the Tier-2 extension in numpy float32.

Test infrastructure.  Written from the specification text (pbrt-v3 7.8 / 7.9.2 as SURVEY.md App. A states it), NOT from
oracle/pbrt_oracle.c: samples are scattered one at a time with the whole footprint of a sample handled as one numpy
expression, pixels live in separate (H, W) planes, and the filters are closed forms over arrays.  Its purpose is to
catch a mistake the C oracle and the CUDA kernels could share, since one author wrote both.  numpy float32 arithmetic
is IEEE, unfused: `a += b * c` rounds the product, then the sum.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
TABLE_WIDTH = 16


# ---------------------------------------------------------------- filters (App. A.2), float64 closed forms

def triangle(x, y, r):
    return np.maximum(0.0, r[0] - np.abs(x)) * np.maximum(0.0, r[1] - np.abs(y))


def gaussian(x, y, r, alpha):
    def g(d, rad):
        return np.maximum(0.0, np.exp(-alpha * d * d) - math.exp(-alpha * rad * rad))
    return g(x, r[0]) * g(y, r[1])


def mitchell_1d(x, B, C):
    x = np.abs(2.0 * x)
    outer = ((-B - 6 * C) * x ** 3 + (6 * B + 30 * C) * x ** 2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) / 6.0
    inner = ((12 - 9 * B - 6 * C) * x ** 3 + (-18 + 12 * B + 6 * C) * x ** 2 + (6 - 2 * B)) / 6.0
    return np.where(x > 1, outer, inner)


def mitchell(x, y, r, B, C):
    return mitchell_1d(x / r[0], B, C) * mitchell_1d(y / r[1], B, C)


def sinc(x):
    x = np.abs(x)
    safe = np.where(x < 1e-5, 1.0, x)
    return np.where(x < 1e-5, 1.0, np.sin(np.pi * safe) / (np.pi * safe))


def lanczos(x, y, r, tau):
    def w(d, rad):
        d = np.abs(d)
        return np.where(d > rad, 0.0, sinc(d) * sinc(d / tau))
    return w(x, r[0]) * w(y, r[1])


def evaluate(name, x, y, r, p0=0.0, p1=0.0):
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if name == "box":
        return np.ones(np.broadcast(x, y).shape)
    if name == "triangle":
        return triangle(x, y, r)
    if name == "gaussian":
        return gaussian(x, y, r, p0)
    if name == "mitchell":
        return mitchell(x, y, r, p0, p1)
    if name == "lanczos":
        return lanczos(x, y, r, p0)
    raise ValueError(name)


def filter_table(name, r, p0=0.0, p1=0.0):
    """table[y*16 + x] = evaluate((x + .5) * r.x / 16, (y + .5) * r.y / 16) — film.rs:113-123, in float64"""
    c = (np.arange(TABLE_WIDTH) + 0.5) / TABLE_WIDTH
    return evaluate(name, (c * r[0])[None, :], (c * r[1])[:, None], r, p0, p1).reshape(-1)


# ---------------------------------------------------------------- film (App. A.1), float32

class FilmNp:
    def __init__(self, res, crop, radius, table, max_lum=math.inf):
        # film.rs:92-101: ceil(res * crop)
        self.x0 = int(math.ceil(f32(res[0]) * f32(crop[0]))); self.y0 = int(math.ceil(f32(res[1]) * f32(crop[1])))
        self.x1 = int(math.ceil(f32(res[0]) * f32(crop[2]))); self.y1 = int(math.ceil(f32(res[1]) * f32(crop[3])))
        self.r = (f32(radius[0]), f32(radius[1]))
        self.inv_r = (f32(1) / self.r[0], f32(1) / self.r[1])
        self.table = np.asarray(table, dtype=f32).reshape(TABLE_WIDTH, TABLE_WIDTH)
        self.max_lum = f32(max_lum)
        h, w = self.y1 - self.y0, self.x1 - self.x0
        self.xyz = np.zeros((3, h, w), dtype=f32)
        self.wsum = np.zeros((h, w), dtype=f32)

    def new_tile(self):
        h, w = self.y1 - self.y0, self.x1 - self.x0
        return np.zeros((3, h, w), dtype=f32), np.zeros((h, w), dtype=f32)

    def add_sample(self, tile, p, L, sw):
        """FilmTile::AddSample on a tile that spans the whole cropped film."""
        rgb, wsum = tile
        L = np.asarray(L, dtype=f32).copy()
        ly = f32(0.212671) * L[0] + f32(0.715160) * L[1] + f32(0.072169) * L[2]
        if ly > self.max_lum:
            L *= self.max_lum / ly
        pd = (f32(p[0]) - f32(0.5), f32(p[1]) - f32(0.5))
        lo_x = max(int(math.ceil(pd[0] - self.r[0])), self.x0); hi_x = min(int(math.floor(pd[0] + self.r[0])) + 1, self.x1)
        lo_y = max(int(math.ceil(pd[1] - self.r[1])), self.y0); hi_y = min(int(math.floor(pd[1] + self.r[1])) + 1, self.y1)
        if hi_x <= lo_x or hi_y <= lo_y:
            return
        def bins(lo, hi, c, inv):
            v = np.abs((np.arange(lo, hi).astype(f32) - c) * inv * f32(TABLE_WIDTH))
            return np.minimum(np.floor(v), TABLE_WIDTH - 1).astype(np.int64)
        w = self.table[bins(lo_y, hi_y, pd[1], self.inv_r[1])[:, None], bins(lo_x, hi_x, pd[0], self.inv_r[0])[None, :]]
        ys, xs = slice(lo_y - self.y0, hi_y - self.y0), slice(lo_x - self.x0, hi_x - self.x0)
        for c in range(3):
            rgb[c, ys, xs] += (L[c] * f32(sw)) * w
        wsum[ys, xs] += w

    def merge(self, tile):
        """Film::merge_film_tile (film.rs:313-326): xyz += to_xyz(rgb) evaluated left to right, weight += weight"""
        rgb, wsum = tile
        m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]], dtype=f32)
        for i in range(3):
            self.xyz[i] += (m[i, 0] * rgb[0] + m[i, 1] * rgb[1]) + m[i, 2] * rgb[2]
        self.wsum += wsum

    def add_samples_pass(self, xy, rgbw):
        tile = self.new_tile()
        for p, l in zip(np.asarray(xy, dtype=f32), np.asarray(rgbw, dtype=f32)):
            self.add_sample(tile, p, l[:3], l[3])
        self.merge(tile)

    def pixels_xyzw(self):
        return np.concatenate([self.xyz, self.wsum[None]], axis=0).transpose(1, 2, 0).reshape(-1, 4)
