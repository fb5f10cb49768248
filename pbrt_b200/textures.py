"""`Texture<T>` trait (src/core/texture.rs:24-40) and `ConstantTexture<T>` (src/textures/constant.rs).

`SurfaceInteraction` is a zero-sized struct in the reference (src/core/interaction.rs:22-23), so a
lookup has no input: the scalar `evaluate` stays on the host exactly as in the reference, and the
throughput path is `evaluate_batch`, a device fill of n lookups.
Checkerboard / imagemap textures do not exist in the reference (src/core/api.rs:914-917) and are
not built here.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional

import numpy as np

from . import _lib
from .runtime import as_pointer


class SurfaceInteraction:
    """src/core/interaction.rs:22-23 — carries nothing."""


class Texture:
    def evaluate(self, si: SurfaceInteraction):
        raise NotImplementedError


class ConstantTexture(Texture):
    """constant.rs:32-154. `value` is a float (T = Float) or an RGB triple (T = Spectrum)."""

    def __init__(self, value):
        if np.ndim(value) == 0:
            self.value = float(np.float32(value))
            self.is_spectrum = False
        else:
            v = np.asarray(value, dtype=np.float32)
            if v.shape != (3,):
                raise ValueError("Spectrum value must have 3 components")
            self.value = v
            self.is_spectrum = True

    @staticmethod
    def new(value) -> "ConstantTexture":
        return ConstantTexture(value)

    def evaluate(self, si: Optional[SurfaceInteraction] = None):
        """constant.rs:139-141: `self.value.clone()`."""
        return self.value.copy() if self.is_spectrum else self.value

    def evaluate_batch(self, n: int, out=None):
        """n lookups on the device: (n,) f32 or (n, 3) f32. `out` may be a host or device buffer."""
        if out is None:
            out = np.empty((n, 3) if self.is_spectrum else (n,), dtype=np.float32)
        ptr, is_dev, keep = as_pointer(out)
        if self.is_spectrum:
            _lib.check(_lib.lib.pbrt_texture_constant_eval_rgb(_lib.f32arr(self.value), int(n), ptr, is_dev))
        else:
            _lib.check(_lib.lib.pbrt_texture_constant_eval_f32(self.value, int(n), ptr, is_dev))
        return out

    def __repr__(self) -> str:  # constant.rs:144-154
        return f"ConstantTexture{{{self.value!r}}}"


def create_constant_float_texture(tex2world=None, tp: Optional[Mapping] = None) -> ConstantTexture:
    """constant.rs:61-68: `value` defaults to 1."""
    return ConstantTexture((tp or {}).get("value", 1.0))


def create_constant_spectrum_texture(tex2world=None, tp: Optional[Mapping] = None) -> ConstantTexture:
    """constant.rs:96-103: `value` defaults to Spectrum::from(1.)."""
    return ConstantTexture((tp or {}).get("value", (1.0, 1.0, 1.0)))


def weight_lut() -> np.ndarray:
    """src/core/mipmap.rs:43-52 — the only implemented piece of MIPMap, evaluated on the device."""
    out = np.empty(128, dtype=np.float32)
    _lib.check(_lib.lib.pbrt_mipmap_weight_lut(out.ctypes.data_as(C.POINTER(C.c_float))))
    return out
