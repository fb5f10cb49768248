"""`Filter` trait (src/core/filter.rs:22-29) and its implementations, host side.

`BoxFilter` is the reference's (src/filters/box.rs).  Triangle / Gaussian / Mitchell /
LanczosSinc are named by the reference's factory (src/core/api.rs:954) but not implemented
there: they are EXTENSIONS with no reference parity, following pbrt-v3 ch. 7.8.

Filters never run on the device.  The film only ever sees a filter through its radius and the
256 `evaluate` calls of `Film::new` (src/core/film.rs:113-123), so a user-defined subclass of
`Filter` written in Python works unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Tuple

from . import _lib
from .geometry import vec2


class Filter:
    """trait Filter (filter.rs:22-29)."""

    def evaluate(self, p) -> float:
        raise NotImplementedError

    def radius(self) -> Tuple[float, float]:
        raise NotImplementedError

    def inv_radius(self) -> Tuple[float, float]:
        raise NotImplementedError


class _NativeFilter(Filter):
    """A filter whose formula lives in libpbrt_b200's host code."""

    def __init__(self, kind: int, radius, p0: float = 0.0, p1: float = 0.0):
        rx, ry = vec2(radius)
        h = C.c_void_p()
        _lib.check(_lib.lib.pbrt_filter_create(kind, rx, ry, p0, p1, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.lib.pbrt_filter_destroy(h)

    def evaluate(self, p) -> float:
        return float(_lib.lib.pbrt_filter_evaluate(self._h, float(p[0]), float(p[1])))

    def radius(self) -> Tuple[float, float]:
        out = (C.c_float * 2)()
        _lib.lib.pbrt_filter_radius(self._h, out)
        return (out[0], out[1])

    def inv_radius(self) -> Tuple[float, float]:
        out = (C.c_float * 2)()
        _lib.lib.pbrt_filter_inv_radius(self._h, out)
        return (out[0], out[1])

    def table(self):
        """The 16x16 table of film.rs:113-123, computed natively (same result as 256 evaluate calls)."""
        import numpy as np

        t = np.empty(256, dtype=np.float32)
        _lib.check(_lib.lib.pbrt_filter_table(self._h, t.ctypes.data_as(C.POINTER(C.c_float))))
        return t


class BoxFilter(_NativeFilter):
    """src/filters/box.rs:30-77."""

    def __init__(self, radius):
        super().__init__(_lib.FILTER_BOX, radius)

    @staticmethod
    def new(radius) -> "BoxFilter":
        return BoxFilter(radius)

    @staticmethod
    def create_box_filter(ps: Optional[Mapping[str, float]] = None) -> "BoxFilter":
        """box.rs:57-61 — `ps` stands in for the ParamSet; xwidth / ywidth default to 0.5."""
        ps = ps or {}
        return BoxFilter((ps.get("xwidth", 0.5), ps.get("ywidth", 0.5)))


class TriangleFilter(_NativeFilter):
    """EXTENSION (no reference parity). pbrt-v3 defaults: radius 2."""

    def __init__(self, radius=(2.0, 2.0)):
        super().__init__(_lib.FILTER_TRIANGLE, radius)


class GaussianFilter(_NativeFilter):
    """EXTENSION (no reference parity). pbrt-v3 defaults: radius 2, alpha 2."""

    def __init__(self, radius=(2.0, 2.0), alpha: float = 2.0):
        super().__init__(_lib.FILTER_GAUSSIAN, radius, alpha)


class MitchellFilter(_NativeFilter):
    """EXTENSION (no reference parity). pbrt-v3 defaults: radius 2, B = C = 1/3."""

    def __init__(self, radius=(2.0, 2.0), b: float = 1.0 / 3.0, c: float = 1.0 / 3.0):
        super().__init__(_lib.FILTER_MITCHELL, radius, b, c)


class LanczosSincFilter(_NativeFilter):
    """EXTENSION (no reference parity). pbrt-v3 defaults: radius 4, tau 3."""

    def __init__(self, radius=(4.0, 4.0), tau: float = 3.0):
        super().__init__(_lib.FILTER_LANCZOS, radius, tau)


def make_filter(name: str, ps: Optional[Mapping[str, float]] = None) -> Filter:
    """src/core/api.rs:951-964. Unknown names are an error there (`exit(1)`); here ValueError."""
    ps = ps or {}
    if name == "box":
        return BoxFilter.create_box_filter(ps)
    if name == "triangle":
        return TriangleFilter((ps.get("xwidth", 2.0), ps.get("ywidth", 2.0)))
    if name == "gaussian":
        return GaussianFilter((ps.get("xwidth", 2.0), ps.get("ywidth", 2.0)), ps.get("alpha", 2.0))
    if name == "mitchell":
        return MitchellFilter((ps.get("xwidth", 2.0), ps.get("ywidth", 2.0)), ps.get("B", 1.0 / 3.0), ps.get("C", 1.0 / 3.0))
    if name == "sinc":
        return LanczosSincFilter((ps.get("xwidth", 4.0), ps.get("ywidth", 4.0)), ps.get("tau", 3.0))
    raise ValueError(f"Filter '{name}' unknown.")
