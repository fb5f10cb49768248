"""Host-side image containers: the file formats either side of the film path.

Mirrors src/core/imageio.rs of wathiede/pbrt (`write_image` :235-283, `read_image` :142-184,
PFM :87-140 / :186-213).  The per-pixel arithmetic of the 8-bit path (gamma + to_byte) runs on
the device in `Film.resolve_rgb8`; this module only packs bytes into PNG / PFM containers, which
is sequential host work (zlib deflate) and stays on the host as it does in the reference (`png`
crate).  `to_byte` below is the host mirror used when a caller hands float data straight to
`write_image`.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Tuple

import numpy as np

from .geometry import Bounds2i, Point2i

f32 = np.float32


def gamma_correct(v: np.ndarray) -> np.ndarray:
    """src/lib.rs:93-99, elementwise in f32."""
    v = np.asarray(v, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        hi = f32(1.055) * np.power(v, f32(1.0 / 2.4), dtype=np.float32) - f32(0.055)
    return np.where(v <= f32(0.0031308), f32(12.92) * v, hi).astype(np.float32)


def to_byte(v: np.ndarray) -> np.ndarray:
    """src/core/imageio.rs:66-68: clamp(255*gamma(v)+0.5, 0, 255) as u8 (NaN -> 0)."""
    c = f32(255.0) * gamma_correct(v) + f32(0.5)
    c = np.where(c < 0, f32(0), np.where(c > 255, f32(255), c))
    c = np.where(np.isnan(c), f32(0), c)
    return c.astype(np.uint8)


def _png_chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png8(name: str, rgb8: np.ndarray, resolution: Tuple[int, int]) -> None:
    """8-bit RGB PNG (colour type 2), as imageio.rs:256-270 configures the encoder."""
    w, h = int(resolution[0]), int(resolution[1])
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8).reshape(h, w * 3)
    raw = np.empty((h, w * 3 + 1), dtype=np.uint8)
    raw[:, 0] = 0  # filter type None
    raw[:, 1:] = rgb8
    data = b"\x89PNG\r\n\x1a\n"
    data += _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
    data += _png_chunk(b"IDAT", zlib.compress(raw.tobytes(), 6))
    data += _png_chunk(b"IEND", b"")
    Path(name).write_bytes(data)


def write_pfm(name: str, rgb: np.ndarray, resolution: Tuple[int, int]) -> None:
    """imageio.rs:186-213: 'PF', rows bottom-to-top, scale -1 = little-endian."""
    w, h = int(resolution[0]), int(resolution[1])
    a = np.ascontiguousarray(rgb, dtype="<f4").reshape(h, w * 3)
    with open(name, "wb") as f:
        f.write(f"PF\n{w} {h}\n-1\n".encode())
        f.write(a[::-1].tobytes())


def write_image(name: str, rgb, output_bounds, total_resolution=None) -> None:
    """imageio.rs:235-283: dispatch on the extension; unknown extensions raise."""
    b = Bounds2i.of(output_bounds)
    res = b.diagonal()
    ext = Path(name).suffix.lower().lstrip(".")
    if ext == "png":
        write_png8(name, to_byte(np.asarray(rgb, dtype=np.float32)), res)
    elif ext == "pfm":
        write_pfm(name, rgb, res)
    elif ext in ("exr", "tga"):
        raise NotImplementedError(f"writing .{ext} files is not implemented")  # imageio.rs:272-273
    else:
        raise ValueError(f"unknown file extension {ext}")


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def read_png8(name: str) -> Tuple[np.ndarray, Point2i]:
    raw = Path(name).read_bytes()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(raw):
        (n,) = struct.unpack(">I", raw[pos : pos + 4])
        tag = raw[pos + 4 : pos + 8]
        body = raw[pos + 8 : pos + 8 + n]
        pos += 12 + n
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or ctype != 2 or interlace != 0:
        raise NotImplementedError("only 8-bit non-interlaced RGB PNG is supported")  # the reference assumes RGB8 too
    data = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).reshape(h, w * 3 + 1)
    out = np.zeros((h, w * 3), dtype=np.uint8)
    bpp = 3
    for y in range(h):
        ft, line = int(data[y, 0]), data[y, 1:].astype(np.int32)
        prev = out[y - 1].astype(np.int32) if y else np.zeros(w * 3, dtype=np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(w * 3, dtype=np.int32)
            for i in range(w * 3):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                pred = a if ft == 1 else ((a + b) >> 1 if ft == 3 else _paeth(a, b, c))
                cur[i] = (line[i] + pred) & 255
        out[y] = cur.astype(np.uint8)
    return out.reshape(h * w, 3), Point2i(w, h)


def read_pfm(name: str) -> Tuple[np.ndarray, Point2i]:
    """imageio.rs:87-140."""
    raw = Path(name).read_bytes()
    pos = 0

    def word():
        nonlocal pos
        while raw[pos : pos + 1] in (b" ", b"\n", b"\t"):
            pos += 1
        s = pos
        while raw[pos : pos + 1] not in (b" ", b"\n", b"\t", b""):
            pos += 1
        w = raw[s:pos].decode()
        pos += 1  # the single whitespace byte that ends the word
        return w

    hdr = word()
    if hdr not in ("PF", "Pf"):
        raise ValueError(f"invalid header '{hdr}'")
    nch = 3 if hdr == "PF" else 1
    w, h, scale = int(word()), int(word()), float(word())
    dt = "<f4" if scale < 0 else ">f4"
    a = np.frombuffer(raw, dtype=dt, count=nch * w * h, offset=pos).reshape(h, w, nch)
    a = (a[::-1].astype(np.float32) * f32(abs(scale))).astype(np.float32)
    if nch == 1:
        a = np.repeat(a, 3, axis=2)
    return a.reshape(h * w, 3), Point2i(w, h)


def read_image(name: str) -> Tuple[np.ndarray, Point2i]:
    """imageio.rs:142-184: (n, 3) f32 RGB and the resolution."""
    ext = Path(name).suffix.lower().lstrip(".")
    if ext == "png":
        rgb8, res = read_png8(name)
        return (rgb8.astype(np.float32) / f32(255.0)).astype(np.float32), res
    if ext == "pfm":
        return read_pfm(name)
    if ext in ("exr", "tga"):
        raise NotImplementedError(f"reading .{ext} files is not implemented")
    raise ValueError(f"unknown file extension {ext}")
