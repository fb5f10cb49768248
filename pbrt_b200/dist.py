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


def route_plan(sample_bounds: Bounds2i, cropped: Bounds2i, radius, nranks: int, src_rows: Tuple[int, int]) -> List[Tuple[int, int]]:
    """Rows of a source's block that each shard needs (own rows + halo); see pbrt_film_route_plan."""
    import ctypes as C

    from . import _lib

    out = (C.c_int32 * (2 * nranks))()
    _lib.check(_lib.lib.pbrt_film_route_plan(_lib.i32x4(sample_bounds.as4()), _lib.i32x4(cropped.as4()),
                                             (C.c_float * 2)(float(radius[0]), float(radius[1])), int(nranks),
                                             (C.c_int32 * 2)(int(src_rows[0]), int(src_rows[1])), out))
    return [(out[2 * g], out[2 * g + 1]) for g in range(nranks)]


def route_samples(xy, rgbw, src_rows: Tuple[int, int], sample_bounds: Bounds2i, spp: int, cropped: Bounds2i, radius,
                  rank: int, nranks: int, group=None):
    """Route pixel-major sample streams to the row shards that own them.

    Every rank passes the block it holds: torch tensors `xy` (n, 2) and `rgbw` (n, 4) with the samples of nominal rows
    [src_rows[0], src_rows[1]) of the global stream over `sample_bounds` (n = rows * width * spp; an empty block is
    fine).  The blocks of all ranks must partition the rows of `sample_bounds`.  Returns (xy, rgbw, bounds): this rank's
    shard stream — its own rows plus the halo, assembled in row order — ready for Film.add_samples_tile(bounds, ...).

    A row is one contiguous run of a pixel-major stream, so "bucket by owning shard" is a slice per destination and
    the duplication of rows within the halo of a shard edge is two overlapping slices: no kernel, no staging copy, one
    group of point-to-point sends (NCCL over NVLink on GPUs, gloo in the CPU tests).  Per-pixel order is unchanged, so a
    routed render is bit-identical to the single film."""
    import torch
    import torch.distributed as dist

    width = sample_bounds.p_max.x - sample_bounds.p_min.x
    per_row = width * spp
    assert xy.shape[0] == (src_rows[1] - src_rows[0]) * per_row and rgbw.shape[0] == xy.shape[0]
    # everybody's block
    mine = torch.tensor([int(src_rows[0]), int(src_rows[1])], dtype=torch.int64, device=xy.device)
    if nranks > 1:
        blocks = [torch.empty_like(mine) for _ in range(nranks)]
        dist.all_gather(blocks, mine, group=group)
        blocks = [(int(b[0]), int(b[1])) for b in blocks]
    else:
        blocks = [(int(src_rows[0]), int(src_rows[1]))]
    need = route_plan(sample_bounds, cropped, radius, nranks, (sample_bounds.p_min.y, sample_bounds.p_max.y))[rank]
    out_rows = need[1] - need[0]
    oxy = torch.empty((out_rows * per_row, 2), dtype=xy.dtype, device=xy.device)
    orgbw = torch.empty((out_rows * per_row, 4), dtype=rgbw.dtype, device=rgbw.device)
    ops, local = [], None
    send = route_plan(sample_bounds, cropped, radius, nranks, src_rows)
    for g in range(nranks):  # what I owe shard g
        a, b = send[g]
        if b <= a:
            continue
        lo, hi = (a - src_rows[0]) * per_row, (b - src_rows[0]) * per_row
        if g == rank:
            local = (a, b, lo, hi)
        else:
            ops += [dist.P2POp(dist.isend, xy[lo:hi], g, group), dist.P2POp(dist.isend, rgbw[lo:hi], g, group)]
    for s in range(nranks):  # what source s owes me
        a, b = route_plan(sample_bounds, cropped, radius, nranks, blocks[s])[rank]
        if b <= a:
            continue
        lo, hi = (a - need[0]) * per_row, (b - need[0]) * per_row
        if s == rank:
            oxy[lo:hi].copy_(xy[local[2]:local[3]])
            orgbw[lo:hi].copy_(rgbw[local[2]:local[3]])
        else:
            ops += [dist.P2POp(dist.irecv, oxy[lo:hi], s, group), dist.P2POp(dist.irecv, orgbw[lo:hi], s, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return oxy, orgbw, Bounds2i.raw(sample_bounds.p_min.x, need[0], sample_bounds.p_max.x, max(need[0], need[1]))


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


class _DeviceArray:
    """A device pointer dressed as __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class FrameExchange:
    """Final assembly of a row-sharded film as ONE kernel per rank over peer memory.

    Every rank owns a full-frame rgb buffer; the buffers are exported through CUDA IPC and mapped by
    every other rank once.  `assemble(film)` then runs `resolve_to_frames_kernel`: it resolves the
    rank's rows (the write_image pixel loop) and stores them directly into all `nranks` frames —
    local HBM for its own, NVLink peer stores for the others — so the all-gather needs no staging
    buffer and no separate collective; a barrier afterwards makes every frame complete.
    `assemble_film_rgb` (resolve, then NCCL all_gather_into_tensor) is the two-step baseline.
    """

    def __init__(self, film, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        from .runtime import DeviceBuffer

        self.film, self.group = film, group
        self.rank, self.nranks = film.rank, film.nranks
        c = film.cropped_pixel_bounds
        self.w, self.h = c.p_max.x - c.p_min.x, c.p_max.y - c.p_min.y
        self.frame = DeviceBuffer(max(self.w * self.h, 1) * 12)
        self.frame.zero()
        handle = (C.c_uint8 * 64)()
        _lib.check(_lib.lib.pbrt_b200_ipc_export(C.c_void_p(self.frame.ptr), handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = torch.empty((self.nranks, 64), dtype=torch.uint8, device="cuda")
        if self.nranks > 1:
            dist.all_gather_into_tensor(allh, mine, group=group)
        else:
            allh[0] = mine
        allh = allh.cpu().numpy()
        self.ptrs, self._opened = [], []
        for r in range(self.nranks):
            if r == self.rank:
                self.ptrs.append(self.frame.ptr)
                continue
            hb = (C.c_uint8 * 64)(*[int(v) for v in allh[r]])
            p = C.c_void_p()
            _lib.check(_lib.lib.pbrt_b200_ipc_import(hb, C.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        self._arr = (C.c_void_p * self.nranks)(*self.ptrs)

    def assemble(self, splat_scale: float = 1.0):
        """Returns the (H, W, 3) f32 frame (a view of this rank's buffer), complete on every rank."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        _lib.check(_lib.lib.pbrt_film_resolve_rgb_to_frames(self.film._h, float(splat_scale), self.nranks, self._arr))
        torch.cuda.synchronize()
        if self.nranks > 1:
            dist.barrier(group=self.group)  # every peer's stores have landed in this rank's frame
        return torch.as_tensor(_DeviceArray(self.frame.ptr, (self.h, self.w, 3)), device="cuda")

    def launch(self, splat_scale: float = 1.0) -> None:
        """Just the kernel (for timing on the device); `assemble` adds the synchronisation."""
        from . import _lib

        _lib.check(_lib.lib.pbrt_film_resolve_rgb_to_frames(self.film._h, float(splat_scale), self.nranks, self._arr))

    def close(self) -> None:
        import ctypes as C

        from . import _lib

        for p in self._opened:
            _lib.lib.pbrt_b200_ipc_close(C.c_void_p(p))
        self._opened = []
        self.frame.free()
