"""Device selection, stream binding and raw device buffers over the C ABI's [UTIL] entry points."""
from __future__ import annotations

import ctypes as C
from typing import Any, Tuple

import numpy as np

from . import _lib


def init(device: int = 0) -> None:
    _lib.check(_lib.lib.pbrt_b200_init(int(device)))


def bind_host_to_device_numa(device: int = 0) -> list[int] | None:
    """Pin the calling process to the CPU cores next to `device` (NVML's ideal CPU affinity for the GPU).

    Host buffers allocated afterwards (first touch) then sit in the NUMA node whose PCIe root the GPU hangs
    off, which is what the host-fed path needs when several ranks stream samples at once: without it every
    rank's uploads cross the socket interconnect.  Returns the core list, or None when NVML cannot tell
    (then nothing is changed).  Host-side plumbing only; no effect on results.
    """
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(device)
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                index = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:  # noqa: BLE001 - best effort: NVML absent, cgroup restrictions, ...
        return None


def set_stream(cuda_stream_ptr: int | None) -> None:
    """Run subsequent work on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream).

    None selects the library's own stream.  A handle of 0 is CUDA's legacy default stream, which
    the C ABI spells cudaStreamLegacy (0x1) because NULL there means "library stream".
    """
    if cuda_stream_ptr is None:
        _lib.check(_lib.lib.pbrt_b200_set_stream(C.c_void_p(0)))
    else:
        _lib.check(_lib.lib.pbrt_b200_set_stream(C.c_void_p(int(cuda_stream_ptr) or 1)))


def synchronize() -> None:
    _lib.check(_lib.lib.pbrt_b200_synchronize())


def launch_count() -> int:
    return int(_lib.lib.pbrt_b200_launch_count())


def overlap_passes(on: bool) -> bool:
    """Let consecutive splat passes overlap whatever buffers they read (include/pbrt_b200.h: the caller guarantees the
    samples of a pass were complete before the previous call on the stream was issued).  Returns the previous setting."""
    return bool(_lib.lib.pbrt_b200_overlap_passes(1 if on else 0))


def device_info() -> dict:
    dev, sm, maj, mnr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    mem = C.c_uint64()
    _lib.check(_lib.lib.pbrt_b200_device_info(C.byref(dev), C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return {"device": dev.value, "sm_count": sm.value, "cc": (maj.value, mnr.value), "hbm_bytes": mem.value}


class DeviceBuffer:
    """A raw device allocation (pbrt_b200_malloc / pbrt_b200_free)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _lib.check(_lib.lib.pbrt_b200_malloc(int(nbytes), C.byref(p)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    @staticmethod
    def from_numpy(a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = DeviceBuffer(a.nbytes)
        if a.nbytes:
            _lib.check(_lib.lib.pbrt_b200_memcpy_h2d(C.c_void_p(b.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return b

    def to_numpy(self, dtype, shape) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        if out.nbytes:
            _lib.check(_lib.lib.pbrt_b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), out.nbytes))
        return out

    def zero(self) -> None:
        _lib.check(_lib.lib.pbrt_b200_memset(C.c_void_p(self.ptr), 0, self.nbytes))

    def free(self) -> None:
        p, self.ptr = self.ptr, None
        if p:
            _lib.lib.pbrt_b200_free(C.c_void_p(p))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory (pbrt_b200_host_alloc) exposed as a numpy array."""

    def __init__(self, dtype, shape):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(_lib.lib.pbrt_b200_host_alloc(n, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self) -> None:
        p, self.ptr = self.ptr, None
        self.array = None
        if p:
            _lib.lib.pbrt_b200_host_free(C.c_void_p(p))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def as_pointer(buf: Any, dtype=np.float32) -> Tuple[C.c_void_p, int, Any]:
    """Return (pointer, is_device, keepalive) for a numpy array, DeviceBuffer, or CUDA torch tensor."""
    if isinstance(buf, DeviceBuffer):
        return C.c_void_p(buf.ptr), 1, buf
    if isinstance(buf, np.ndarray):
        a = np.ascontiguousarray(buf, dtype=dtype)
        return a.ctypes.data_as(C.c_void_p), 0, a
    if hasattr(buf, "data_ptr") and hasattr(buf, "is_cuda"):  # torch.Tensor
        if not buf.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(buf.data_ptr()), (1 if buf.is_cuda else 0), buf
    a = np.ascontiguousarray(np.asarray(buf, dtype=dtype))
    return a.ctypes.data_as(C.c_void_p), 0, a
