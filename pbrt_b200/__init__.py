"""pbrt_b200 — the film-reconstruction and texture-evaluation path of wathiede/pbrt on B200.

Host-side mirror of the reference's `Film`, `FilmTile`, `Filter`, `BoxFilter`, `Texture` and
`ConstantTexture` over the C ABI in include/pbrt_b200.h (libpbrt_b200.so: hand-written CUDA for
sm_100a).  There is no CPU fallback: importing fails without the built library, and every
compute call fails without a B200.
"""
from . import _lib  # noqa: F401  (raises ImportError if the CUDA library is not built)
from ._lib import PbrtError
from .film import FILTER_TABLE_WIDTH, SPLAT_ATOMIC, SPLAT_EXACT, SPLAT_FMA, Film, FilmTile, FilmTilePixel, filter_table
from .filters import (BoxFilter, Filter, GaussianFilter, LanczosSincFilter, MitchellFilter, TriangleFilter,
                      make_filter)
from .geometry import Bounds2f, Bounds2i, Point2i
from .runtime import (DeviceBuffer, PinnedBuffer, bind_host_to_device_numa, device_info, init, launch_count, overlap_passes, set_stream,
                      synchronize)
from .textures import (ConstantTexture, SurfaceInteraction, Texture, create_constant_float_texture,
                       create_constant_spectrum_texture, weight_lut)

__all__ = [
    "PbrtError", "Film", "FilmTile", "FilmTilePixel", "filter_table", "FILTER_TABLE_WIDTH",
    "SPLAT_EXACT", "SPLAT_FMA", "SPLAT_ATOMIC",
    "Filter", "BoxFilter", "TriangleFilter", "GaussianFilter", "MitchellFilter", "LanczosSincFilter", "make_filter",
    "Bounds2f", "Bounds2i", "Point2i",
    "DeviceBuffer", "PinnedBuffer", "bind_host_to_device_numa", "device_info", "init", "launch_count", "overlap_passes", "set_stream", "synchronize",
    "Texture", "ConstantTexture", "SurfaceInteraction", "create_constant_float_texture",
    "create_constant_spectrum_texture", "weight_lut",
]
