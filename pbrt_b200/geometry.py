"""The 2-D integer/float bounds the film path needs, mirroring the reference's semantics.

Follows src/core/geometry/bounds.rs and point.rs of wathiede/pbrt (cited per method).  Only the
2-D pieces on the film path exist here; 3-D geometry is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, Sequence, Tuple

import numpy as np

f32 = np.float32


@dataclass(frozen=True)
class Point2i:
    x: int
    y: int

    @staticmethod
    def of(p) -> "Point2i":
        if isinstance(p, Point2i):
            return p
        return Point2i(int(p[0]), int(p[1]))

    def __iter__(self):
        yield self.x
        yield self.y


@dataclass(frozen=True)
class Bounds2i:
    """`Bounds2<isize>`: p_max is exclusive. Fields may be inverted (see `intersect`)."""

    p_min: Point2i
    p_max: Point2i

    @staticmethod
    def of(b) -> "Bounds2i":
        """`Bounds2i::from([[x0,y0],[x1,y1]])` — sorts each axis (bounds.rs:119-130)."""
        if isinstance(b, Bounds2i):
            return b
        if len(b) == 4:  # flat {x0, y0, x1, y1} as the C ABI passes bounds: taken as is
            return Bounds2i.raw(*b)
        (ax, ay), (bx, by) = b
        return Bounds2i(Point2i(min(ax, bx), min(ay, by)), Point2i(max(ax, bx), max(ay, by)))

    @staticmethod
    def raw(x0: int, y0: int, x1: int, y1: int) -> "Bounds2i":
        """Construct without sorting, as `Bounds2 { p_min, p_max }` does."""
        return Bounds2i(Point2i(int(x0), int(y0)), Point2i(int(x1), int(y1)))

    def as4(self) -> Tuple[int, int, int, int]:
        return (self.p_min.x, self.p_min.y, self.p_max.x, self.p_max.y)

    def diagonal(self) -> Tuple[int, int]:
        return (self.p_max.x - self.p_min.x, self.p_max.y - self.p_min.y)

    def area(self) -> int:
        """bounds.rs:195-198 — a plain product, positive for a doubly inverted box."""
        dx, dy = self.diagonal()
        return dx * dy

    def inside_exclusive(self, p) -> bool:
        """bounds.rs:210-212."""
        p = Point2i.of(p)
        return self.p_min.x <= p.x < self.p_max.x and self.p_min.y <= p.y < self.p_max.y

    @staticmethod
    def intersect(b1: "Bounds2i", b2: "Bounds2i") -> "Bounds2i":
        """bounds.rs:244-252 — the result is deliberately not re-sorted."""
        return Bounds2i(
            Point2i(max(b1.p_min.x, b2.p_min.x), max(b1.p_min.y, b2.p_min.y)),
            Point2i(min(b1.p_max.x, b2.p_max.x), min(b1.p_max.y, b2.p_max.y)),
        )

    def iter(self) -> Iterator[Point2i]:
        """bounds.rs:284-288 — row-major, y outer; an inverted range yields nothing."""
        for y in range(self.p_min.y, self.p_max.y):
            for x in range(self.p_min.x, self.p_max.x):
                yield Point2i(x, y)


@dataclass(frozen=True)
class Bounds2f:
    p_min: Tuple[float, float]
    p_max: Tuple[float, float]

    @staticmethod
    def of(b) -> "Bounds2f":
        if isinstance(b, Bounds2f):
            return b
        (ax, ay), (bx, by) = b
        ax, ay, bx, by = f32(ax), f32(ay), f32(bx), f32(by)
        return Bounds2f((float(min(ax, bx)), float(min(ay, by))), (float(max(ax, bx)), float(max(ay, by))))

    def as4(self) -> Tuple[float, float, float, float]:
        return (self.p_min[0], self.p_min[1], self.p_max[0], self.p_max[1])


def vec2(v: Sequence[float]) -> Tuple[float, float]:
    """A Vector2f / Point2f as a pair of f32-rounded Python floats."""
    return (float(f32(v[0])), float(f32(v[1])))
