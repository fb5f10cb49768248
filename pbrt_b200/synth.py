"""Synthetic workloads of SURVEY.md App. C, generated in HBM with the reference's PCG32 (rng.rs)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .geometry import Bounds2i
from .runtime import DeviceBuffer


def samples(bounds, spp: int, seed: int = 1, index_bounds=None) -> Tuple[DeviceBuffer, DeviceBuffer, int]:
    """Stratified pixel-major sample stream over `bounds`: (xy, rgbw) device buffers and the count."""
    b = Bounds2i.of(bounds)
    n = max(b.area(), 0) * spp
    xy, rgbw = DeviceBuffer(max(n, 1) * 8), DeviceBuffer(max(n, 1) * 16)
    ib = _lib.i32x4(Bounds2i.of(index_bounds).as4()) if index_bounds is not None else None
    _lib.check(_lib.lib.pbrt_synth_samples(_lib.i32x4(b.as4()), ib, int(spp), int(seed), C.c_void_p(xy.ptr), C.c_void_p(rgbw.ptr)))
    return xy, rgbw, n


def tiles(counts, seed: int = 1) -> Tuple[DeviceBuffer, np.ndarray, int]:
    """Tile fill for the merge workload: device rgbw for tiles with the given pixel counts."""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    total = int(counts.sum())
    buf = DeviceBuffer(max(total, 1) * 16)
    _lib.check(
        _lib.lib.pbrt_synth_tiles(
            len(counts), offsets.ctypes.data_as(C.POINTER(C.c_int64)), counts.ctypes.data_as(C.POINTER(C.c_int64)),
            int(seed), C.c_void_p(buf.ptr), total,
        )
    )
    return buf, offsets, total
