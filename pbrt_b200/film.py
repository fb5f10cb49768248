"""`Film` / `FilmTile` (src/core/film.rs) over the C ABI: same names, arguments and error behaviour.

The pixel storage lives in HBM behind a `PbrtFilm*`; the bounds arithmetic and both hot loops
(`merge_film_tile`, `write_image`) run in libpbrt_b200.  Where the reference panics
(`unwrap`, `debug_assert!`) this raises `PbrtError` with code E_RANGE.

Methods marked EXTENSION have no implementation in the reference (`unimplemented!()` or absent)
and therefore no reference parity; they follow pbrt-v3.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, imageio
from .filters import Filter, _NativeFilter
from .geometry import Bounds2f, Bounds2i, Point2i, f32
from .runtime import as_pointer

FILTER_TABLE_WIDTH = 16  # film.rs:34

SPLAT_EXACT, SPLAT_FMA, SPLAT_ATOMIC = _lib.SPLAT_EXACT, _lib.SPLAT_FMA, _lib.SPLAT_ATOMIC


def filter_table(filt: Filter) -> np.ndarray:
    """film.rs:113-123: table[y*16+x] = filter.evaluate(((x+.5)*r.x)/16, ((y+.5)*r.y)/16) in f32."""
    if isinstance(filt, _NativeFilter) and type(filt).evaluate is _NativeFilter.evaluate:
        return filt.table()
    w = f32(FILTER_TABLE_WIDTH)
    rx, ry = (f32(v) for v in filt.radius())
    t = np.empty(256, dtype=np.float32)
    for y in range(FILTER_TABLE_WIDTH):
        for x in range(FILTER_TABLE_WIDTH):
            px = (f32(x) + f32(0.5)) * rx / w
            py = (f32(y) + f32(0.5)) * ry / w
            t[y * FILTER_TABLE_WIDTH + x] = f32(filt.evaluate((float(px), float(py))))
    return t


def _pixel_major_spp(xy: np.ndarray, sb: Optional[Bounds2i]) -> int:
    """spp if `xy` is a pixel-major stream over `sb` with the same count in every pixel, else 0."""
    if sb is None or len(xy) == 0:
        return 0
    w, h = sb.p_max.x - sb.p_min.x, sb.p_max.y - sb.p_min.y
    if w <= 0 or h <= 0 or len(xy) % (w * h):
        return 0
    spp = len(xy) // (w * h)
    px = np.repeat(np.tile(np.arange(sb.p_min.x, sb.p_max.x), h), spp)
    py = np.repeat(np.repeat(np.arange(sb.p_min.y, sb.p_max.y), w), spp)
    ok = (xy[:, 0] >= px) & (xy[:, 0] <= px + 1) & (xy[:, 1] >= py) & (xy[:, 1] <= py + 1)
    return spp if bool(ok.all()) else 0


class FilmTilePixel:
    """film.rs:39-42 — a view of one pixel of a tile's buffer."""

    __slots__ = ("_row",)

    def __init__(self, row: np.ndarray):
        self._row = row

    @property
    def contrib_sum(self) -> np.ndarray:
        return self._row[:3]

    @contrib_sum.setter
    def contrib_sum(self, rgb) -> None:
        self._row[:3] = rgb

    @property
    def filter_weight_sum(self) -> float:
        return float(self._row[3])

    @filter_weight_sum.setter
    def filter_weight_sum(self, w: float) -> None:
        self._row[3] = w


class FilmTile:
    """film.rs:428-489. `pixels` is the Vec<FilmTilePixel>: (pixel_count, 4) f32 {rgb, weight}."""

    def __init__(self, film: "Film", pixel_bounds: Bounds2i, pixel_count: int):
        self._film = film
        self.pixel_bounds = pixel_bounds
        self.pixels = np.zeros((pixel_count, 4), dtype=np.float32)  # FilmTilePixel::default()
        self._samples_xy: List[Tuple[float, float]] = []
        self._samples_rgbw: List[Tuple[float, float, float, float]] = []
        self._sample_bounds: Optional[Bounds2i] = None

    def get_pixel_bounds(self) -> Bounds2i:
        return self.pixel_bounds

    def pixel_offset(self, p) -> int:
        """film.rs:465-476 — panics (here: PbrtError E_RANGE) outside the tile."""
        p = Point2i.of(p)
        if not self.pixel_bounds.inside_exclusive(p):
            raise _lib.PbrtError(_lib.E_RANGE, f"p [{p.x}, {p.y}] outside {self.pixel_bounds.as4()}")
        width = self.pixel_bounds.p_max.x - self.pixel_bounds.p_min.x
        return (p.x - self.pixel_bounds.p_min.x) + (p.y - self.pixel_bounds.p_min.y) * width

    def get_pixel(self, p) -> FilmTilePixel:
        return FilmTilePixel(self.pixels[self.pixel_offset(p)])

    def get_pixel_mut(self, p) -> FilmTilePixel:
        return FilmTilePixel(self.pixels[self.pixel_offset(p)])

    def add_sample(self, p_film, L, sample_weight: float = 1.0) -> None:
        """EXTENSION (pbrt-v3 FilmTile::AddSample). Buffered on the host; splatted on merge."""
        self._samples_xy.append((float(p_film[0]), float(p_film[1])))
        self._samples_rgbw.append((float(L[0]), float(L[1]), float(L[2]), float(sample_weight)))


class Film:
    """film.rs:59-76. Construct with `Film.new(...)` like the reference."""

    def __init__(self):
        raise TypeError("use Film.new(...)")

    @classmethod
    def new(
        cls,
        resolution,
        crop_window,
        filter: Filter,
        diagonal_mm: float,
        filename: str,
        scale: float,
        max_sample_luminance: float,
        *,
        rank: int = 0,
        nranks: int = 1,
    ) -> "Film":
        """film.rs:82-137. `rank`/`nranks` (keyword-only, not in the reference) select a row shard."""
        self = object.__new__(cls)
        self.full_resolution = Point2i.of(resolution)
        self._crop_window = Bounds2f.of(crop_window)
        self.filter = filter
        self.diagonal_m = float(f32(diagonal_mm) * f32(0.001))
        self.filename = filename
        self.scale = float(scale)
        self.max_sample_luminance = float(max_sample_luminance)
        self.filter_table = filter_table(filter)
        self._h = C.c_void_p()
        crop = _lib.f32arr(self._crop_window.as4())
        rad = _lib.f32arr(filter.radius())
        tab = self.filter_table.ctypes.data_as(C.POINTER(C.c_float))
        _lib.check(
            _lib.lib.pbrt_film_create_sharded(
                self.full_resolution.x, self.full_resolution.y, crop, rad, tab, float(diagonal_mm), float(scale),
                float(max_sample_luminance), int(rank), int(nranks), C.byref(self._h),
            )
        )
        b = (C.c_int32 * 4)()
        _lib.check(_lib.lib.pbrt_film_cropped_pixel_bounds(self._h, b))
        self.cropped_pixel_bounds = Bounds2i.raw(*b)
        _lib.check(_lib.lib.pbrt_film_owned_pixel_bounds(self._h, b))
        self.owned_pixel_bounds = Bounds2i.raw(*b)
        self.rank, self.nranks = int(rank), int(nranks)
        return self

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.lib.pbrt_film_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ bounds (film.rs:166-281)
    def get_sample_bounds(self) -> Bounds2i:
        b = (C.c_int32 * 4)()
        _lib.check(_lib.lib.pbrt_film_get_sample_bounds(self._h, b))
        return Bounds2i.raw(*b)

    def get_physical_extent(self) -> Bounds2f:
        e = (C.c_float * 4)()
        _lib.check(_lib.lib.pbrt_film_get_physical_extent(self._h, e))
        return Bounds2f((e[0], e[1]), (e[2], e[3]))

    def _tile_bounds(self, sample_bounds) -> Tuple[Bounds2i, int]:
        sb = Bounds2i.of(sample_bounds)
        out = (C.c_int32 * 4)()
        cnt = C.c_int64()
        _lib.check(_lib.lib.pbrt_film_tile_bounds(self._h, _lib.i32x4(sb.as4()), out, C.byref(cnt)))
        return Bounds2i.raw(*out), cnt.value

    def get_film_tile(self, sample_bounds) -> FilmTile:
        tb, cnt = self._tile_bounds(sample_bounds)
        t = FilmTile(self, tb, cnt)
        t._sample_bounds = Bounds2i.of(sample_bounds)
        return t

    # ------------------------------------------------------------------ merge (film.rs:313-326)
    def merge_film_tile(self, tile: FilmTile) -> None:
        """Consumes `tile` as the reference does (it takes the tile by value)."""
        if tile._samples_xy:  # EXTENSION: samples recorded through FilmTile.add_sample
            xy = np.asarray(tile._samples_xy, dtype=np.float32)
            rgbw = np.asarray(tile._samples_rgbw, dtype=np.float32)
            spp = _pixel_major_spp(xy, tile._sample_bounds)
            if spp:   # the order a renderer produces: the exact, order-preserving kernel applies
                self.add_samples_tile(tile._sample_bounds, spp, xy, rgbw, SPLAT_EXACT)
            else:     # any other order: scatter (order of additions not fixed)
                self.add_samples(tile._sample_bounds, xy, rgbw)
            tile._samples_xy, tile._samples_rgbw = [], []
            # a broken add_sample contract (sample outside its nominal pixel, non-finite radiance) surfaces here, where the
            # reference's debug_assert!s would have fired, not at some later check()
            self.check()
        _lib.check(
            _lib.lib.pbrt_film_merge_tile(
                self._h, _lib.i32x4(tile.pixel_bounds.as4()), tile.pixels.ctypes.data_as(C.c_void_p), 0
            )
        )
        tile.pixels = None

    def merge_film_tiles(self, tiles: Sequence[FilmTile]) -> None:
        """All tiles in one launch; same result as merging them one by one in order."""
        tiles = list(tiles)
        if not tiles:
            return
        bounds = np.asarray([t.pixel_bounds.as4() for t in tiles], dtype=np.int32)
        counts = np.asarray([len(t.pixels) for t in tiles], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        rgbw = np.concatenate([t.pixels for t in tiles], axis=0) if counts.sum() else np.zeros((0, 4), np.float32)
        self.merge_tiles_raw(bounds, offsets, rgbw)
        for t in tiles:
            t.pixels = None

    def merge_tile_raw(self, tile_bounds, rgbw) -> None:
        """One tile from a host or device buffer of (pixels, 4) f32."""
        tb = Bounds2i.of(tile_bounds) if not isinstance(tile_bounds, Bounds2i) else tile_bounds
        ptr, is_dev, keep = as_pointer(rgbw)
        _lib.check(_lib.lib.pbrt_film_merge_tile(self._h, _lib.i32x4(tb.as4()), ptr, is_dev))

    def merge_tiles_raw(self, bounds: np.ndarray, offsets: np.ndarray, rgbw, total_pixels: Optional[int] = None) -> None:
        bounds = np.ascontiguousarray(bounds, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ptr, is_dev, keep = as_pointer(rgbw)
        if total_pixels is None:
            total_pixels = int(np.asarray(keep).size // 4) if not is_dev else int(keep.nbytes // 16)
        _lib.check(
            _lib.lib.pbrt_film_merge_tiles(
                self._h, len(bounds), bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                offsets.ctypes.data_as(C.POINTER(C.c_int64)), ptr, int(total_pixels), is_dev,
            )
        )

    # ------------------------------------------------------------------ EXTENSION: sample splatting
    def add_samples_tile(self, sample_bounds, spp: int, xy, rgbw, mode: int = SPLAT_EXACT, pinned_async: bool = False) -> None:
        """get_film_tile(sample_bounds) -> add_sample for each sample -> merge_film_tile, on the device.

        Samples are pixel-major over `sample_bounds`, `spp` per pixel, each inside its pixel.
        `pinned_async`: xy / rgbw are page-locked host arrays (PinnedBuffer.array); the upload is enqueued on
        the copy stream and overlaps the previous call's kernels; keep the arrays untouched until synchronize().
        """
        sb = Bounds2i.of(sample_bounds)
        pxy, dev_a, k1 = as_pointer(xy)
        prgbw, dev_b, k2 = as_pointer(rgbw)
        if dev_a != dev_b:
            raise ValueError("xy and rgbw must both be host or both be device buffers")
        if pinned_async:
            if dev_a:
                raise ValueError("pinned_async is for host arrays")
            dev_a = 2  # PBRT_MEM_PINNED_ASYNC
        _lib.check(_lib.lib.pbrt_film_add_samples_tile(self._h, _lib.i32x4(sb.as4()), int(spp), pxy, prgbw, dev_a, int(mode)))

    def add_samples_tile_rgb(self, sample_bounds, spp: int, xy, rgb, sample_weight=None, mode: int = SPLAT_EXACT,
                             pinned_async: bool = False) -> None:
        """add_samples_tile with the radiance as separate streams: `rgb` (n, 3) and `sample_weight` (n,) or None
        for all-ones (then no weight stream is transferred).  Bit-identical to the interleaved form."""
        sb = Bounds2i.of(sample_bounds)
        pxy, dev_a, k1 = as_pointer(xy)
        prgb, dev_b, k2 = as_pointer(rgb)
        psw, dev_c, k3 = as_pointer(sample_weight) if sample_weight is not None else (None, dev_a, None)
        if not (dev_a == dev_b == dev_c):
            raise ValueError("xy, rgb and sample_weight must all be host or all be device buffers")
        if pinned_async:
            if dev_a:
                raise ValueError("pinned_async is for host arrays")
            dev_a = 2  # PBRT_MEM_PINNED_ASYNC
        _lib.check(_lib.lib.pbrt_film_add_samples_tile_rgb(
            self._h, _lib.i32x4(sb.as4()), int(spp), pxy, prgb, psw, dev_a, int(mode)))

    def add_samples_tiles(self, sample_bounds, spp: int, xy, rgbw, sample_offsets=None, mode: int = SPLAT_EXACT) -> None:
        """Many tiles in one call: tile i has sample bounds `sample_bounds[i]` (x0, y0, x1, y1) and its pixel-major
        samples start at `sample_offsets[i]` (default: tiles packed back to back).  Same result as calling
        add_samples_tile for each tile in order."""
        if isinstance(sample_bounds, np.ndarray) and sample_bounds.ndim == 2 and sample_bounds.shape[1] == 4:
            sbs = np.ascontiguousarray(sample_bounds, dtype=np.int32)  # already flat {x0, y0, x1, y1} rows
        else:
            sbs = np.ascontiguousarray([Bounds2i.of(b).as4() for b in sample_bounds], dtype=np.int32)
        if sample_offsets is None:
            w = np.maximum(sbs[:, 2].astype(np.int64) - sbs[:, 0], 0)
            h = np.maximum(sbs[:, 3].astype(np.int64) - sbs[:, 1], 0)
            counts = w * h * int(spp)
            sample_offsets = np.concatenate([[0], np.cumsum(counts)[:-1]])
        offs = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        pxy, dev_a, k1 = as_pointer(xy)
        prgbw, dev_b, k2 = as_pointer(rgbw)
        if dev_a != dev_b:
            raise ValueError("xy and rgbw must both be host or both be device buffers")
        total = int((np.asarray(k1).size // 2) if not dev_a else (k1.nbytes // 8))
        _lib.check(
            _lib.lib.pbrt_film_add_samples_tiles(
                self._h, len(sbs), sbs.ctypes.data_as(C.POINTER(C.c_int32)), offs.ctypes.data_as(C.POINTER(C.c_int64)),
                int(spp), pxy, prgbw, total, dev_a, int(mode),
            )
        )

    def add_samples(self, sample_bounds, xy, rgbw) -> None:
        """Samples in any order / position (global-atomic scatter; order of additions not fixed)."""
        sb = Bounds2i.of(sample_bounds)
        pxy, dev_a, k1 = as_pointer(xy)
        prgbw, dev_b, k2 = as_pointer(rgbw)
        if dev_a != dev_b:
            raise ValueError("xy and rgbw must both be host or both be device buffers")
        n = (np.asarray(k1).size // 2) if not dev_a else (k1.nbytes // 8)
        _lib.check(_lib.lib.pbrt_film_add_samples(self._h, _lib.i32x4(sb.as4()), int(n), pxy, prgbw, dev_a))

    def check(self) -> None:
        """Raise the sticky asynchronous error of the film's kernels, if any."""
        _lib.check(_lib.lib.pbrt_film_check(self._h))

    # ------------------------------------------------------------------ EXTENSION: film.rs:329-336, :386-388
    def set_image(self, img) -> None:
        ptr, is_dev, keep = as_pointer(img)
        _lib.check(_lib.lib.pbrt_film_set_image(self._h, ptr, is_dev))

    def add_splat(self, p, v) -> None:
        self.add_splats(np.asarray([p], dtype=np.float32), np.asarray([v], dtype=np.float32))

    def add_splats(self, xy, rgb) -> None:
        pxy, dev_a, k1 = as_pointer(xy)
        prgb, dev_b, k2 = as_pointer(rgb)
        n = (np.asarray(k1).size // 2) if not dev_a else (k1.nbytes // 8)
        _lib.check(_lib.lib.pbrt_film_add_splats(self._h, int(n), pxy, prgb, dev_a))

    def clear(self) -> None:
        _lib.check(_lib.lib.pbrt_film_clear(self._h))

    # ------------------------------------------------------------------ write_image (film.rs:340-383)
    def resolve_rgb(self, splat_scale: float = 1.0, out=None, pinned_async: bool = False) -> np.ndarray:
        """The rgb buffer `write_image` builds (film.rs:342-372): (owned pixels, 3) f32.

        `pinned_async`: `out` is a page-locked host array and the read-back is only enqueued (synchronize() completes it).
        """
        n = max(self.owned_pixel_bounds.area(), 0)
        if out is None:
            out = np.empty((n, 3), dtype=np.float32)
        ptr, is_dev, keep = as_pointer(out)
        if pinned_async and not is_dev:
            is_dev = 2  # PBRT_MEM_PINNED_ASYNC
        _lib.check(_lib.lib.pbrt_film_resolve_rgb(self._h, float(splat_scale), ptr, is_dev))
        return out

    def resolve_rgb8(self, splat_scale: float = 1.0, out=None) -> np.ndarray:
        """The same, fused with imageio's to_byte (imageio.rs:66-68): (owned pixels, 3) u8."""
        n = max(self.owned_pixel_bounds.area(), 0)
        if out is None:
            out = np.empty((n, 3), dtype=np.uint8)
        ptr, is_dev, keep = as_pointer(out, dtype=np.uint8)
        _lib.check(_lib.lib.pbrt_film_resolve_rgb8(self._h, float(splat_scale), ptr, is_dev))
        return out

    def write_image(self, splat_scale: float = 1.0) -> None:
        """film.rs:340-383: resolve, then imageio::write_image by file extension."""
        b = self.owned_pixel_bounds
        if self.filename.lower().endswith(".png"):
            rgb8 = self.resolve_rgb8(splat_scale)
            imageio.write_png8(self.filename, rgb8, b.diagonal())
        else:
            rgb = self.resolve_rgb(splat_scale)
            imageio.write_image(self.filename, rgb.reshape(-1), b, self.full_resolution)

    # ------------------------------------------------------------------ read-back
    def get_pixel_xyz(self, p) -> Tuple[float, float, float]:
        """film.rs:405-410."""
        p = Point2i.of(p)
        out = (C.c_float * 3)()
        _lib.check(_lib.lib.pbrt_film_get_pixel_xyz(self._h, p.x, p.y, out))
        return (out[0], out[1], out[2])

    def read_pixels(self) -> np.ndarray:
        """All owned pixels as the reference's `Pixel` (film.rs:47-55): (n, 7) f32."""
        n = max(self.owned_pixel_bounds.area(), 0)
        out = np.empty((n, 7), dtype=np.float32)
        if n:
            _lib.check(_lib.lib.pbrt_film_read_pixels(self._h, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def device_buffers(self) -> Tuple[int, int, int]:
        """(xyzw device pointer, splat device pointer, owned pixel count) for collective plumbing."""
        a, b, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib.pbrt_film_device_buffers(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return (a.value or 0, b.value or 0, n.value)
