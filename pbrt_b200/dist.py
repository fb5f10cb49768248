"""Row sharding of the film across the GPUs of one box.

The film partitions by pixel rows: rank g of G owns rows [y0 + g*H/G, y0 + (g+1)*H/G) of the
cropped pixel bounds, which is exactly a `Film` whose clip rectangle is that row block.  A sample
belongs to every shard its footprint can reach, so samples within h = floor(r.y + .5) rows of a
shard edge are processed by both neighbours and each clips; no cross-GPU reduction exists and the
per-pixel order of additions is unchanged.  The only exchange is the final assembly: one
all-gather of the resolved row blocks (NCCL over NVLink on GPUs; gloo on CPU in the tests).

torch is used for the process group only; it is imported lazily so the rest of the package does
not depend on it.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

from .geometry import Bounds2i


def shard_rows(cropped: Bounds2i, rank: int, nranks: int) -> Tuple[int, int]:
    """Rows owned by `rank`; the same split pbrt_film_create_sharded makes."""
    y0, h = cropped.p_min.y, cropped.p_max.y - cropped.p_min.y
    return (y0 + h * rank // nranks, y0 + h * (rank + 1) // nranks)


def halo_rows(radius_y: float) -> int:
    """A sample in nominal pixel row n reaches rows n-h .. n+h, h = floor(r.y + .5)."""
    return int(math.floor(radius_y + 0.5))


def shard_sample_bounds(sample_bounds: Bounds2i, owned_rows: Tuple[int, int], radius_y: float) -> Bounds2i:
    """The nominal sample rows a shard must see: its own rows plus the halo, within `sample_bounds`."""
    h = halo_rows(radius_y)
    y0 = max(sample_bounds.p_min.y, owned_rows[0] - h)
    y1 = min(sample_bounds.p_max.y, owned_rows[1] + h)
    return Bounds2i.raw(sample_bounds.p_min.x, y0, sample_bounds.p_max.x, max(y0, y1))


def all_rows(cropped: Bounds2i, nranks: int) -> List[Tuple[int, int]]:
    return [shard_rows(cropped, r, nranks) for r in range(nranks)]


def allgather_rows(local, cropped: Bounds2i, rank: int, nranks: int, group=None):
    """Assemble per-rank row blocks into the full frame on every rank.

    `local` is a torch tensor of shape (owned_rows, width, C) on the device the process group
    communicates on.  Row counts may differ by one between ranks; blocks are padded to the
    tallest so that a single all_gather_into_tensor moves everything.
    """
    import torch
    import torch.distributed as dist

    rows = all_rows(cropped, nranks)
    width = cropped.p_max.x - cropped.p_min.x
    hmax = max(b - a for a, b in rows)
    assert local.shape[0] == rows[rank][1] - rows[rank][0] and local.shape[1] == width
    if nranks == 1:
        return local.clone()
    send = local
    if local.shape[0] != hmax:
        send = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    recv = torch.empty((nranks * hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(b - a == hmax for a, b in rows):
        return recv
    parts = [recv[r * hmax : r * hmax + (rows[r][1] - rows[r][0])] for r in range(nranks)]
    return torch.cat(parts, dim=0)


def assemble_film_rgb(film, splat_scale: float = 1.0, group=None):
    """resolve this rank's rows on its GPU, then all-gather: the full (H, W, 3) f32 frame on every rank."""
    import torch

    ob = film.owned_pixel_bounds
    w, h = ob.p_max.x - ob.p_min.x, ob.p_max.y - ob.p_min.y
    local = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    film.resolve_rgb(splat_scale, out=local)
    return allgather_rows(local, film.cropped_pixel_bounds, film.rank, film.nranks, group)
