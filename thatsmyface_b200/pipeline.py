"""Host <-> B200 pipeline around the fused kernels.

The reference's embed page walks the uploaded images one at a time
(``internal_pages/embed_watermark_page.py:492-558``).  Here a batch of host
images goes through the C ABI's host-buffer context (``tmf_ctx_*`` in
``include/tmf_wm.h``): the batch is cut into chunks and H2D copy, fused kernel and
D2H copy of successive chunks overlap on three CUDA streams with ``depth`` device
slots.  With several devices the batch is split purely by image (contiguous
ranges, one context per device, no collective, no peer traffic); every context is
enqueued before any is synchronised, so the devices run concurrently from this
one host thread.

The pipeline itself is native (C++/CUDA runtime inside libtmfwm.so); this module
only validates shapes, hands over host pointers and keeps one context per
(thread, device).  PyTorch is used for pinned host memory and nothing else.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .constants import ALPHA, BLOCK_SIZE

DEFAULT_CHUNK_BYTES = 96 << 20
DEFAULT_DEPTH = 2


class HostPipeline:
    """One ``tmf_ctx`` (streams + device slots) on one device.  Not thread-safe; use one
    per thread (``context_for`` does that)."""

    def __init__(self, device: int, chunk_bytes: int = DEFAULT_CHUNK_BYTES, depth: int = DEFAULT_DEPTH):
        self._lib = _lib.load()
        self.device, self.chunk_bytes, self.depth = int(device), int(chunk_bytes), int(depth)
        h = C.c_void_p()
        _lib.check(self._lib.tmf_ctx_create(C.byref(h), self.device, self.chunk_bytes, self.depth))
        self._h = h

    def embed_async(self, src, dst, n, h, w, wm, wm_shared, alpha, block_size, mode):
        _lib.check(self._lib.tmf_ctx_embed_host_async(self._h, src, dst, n, h, w, wm, 1 if wm_shared else 0,
                                                      float(alpha), int(block_size), int(mode)))

    def extract_async(self, src_a, src_b, dst, n, h, w, alpha, block_size, mode):
        _lib.check(self._lib.tmf_ctx_extract_host_async(self._h, src_a, src_b, dst, n, h, w, float(alpha),
                                                        int(block_size), int(mode)))

    def synchronize(self):
        _lib.check(self._lib.tmf_ctx_synchronize(self._h))

    def stats(self, reset=False):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        _lib.check(self._lib.tmf_ctx_stats(self._h, C.byref(a), C.byref(b), C.byref(c), 1 if reset else 0))
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.tmf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):  # best effort; contexts normally live as long as the thread
        try:
            self.close()
        except Exception:
            pass


class pinned:
    """Context manager that page-locks NumPy arrays for the duration of a batch call, so the
    pipeline's copies are truly asynchronous (tmf_pin_host / tmf_unpin_host)::

        with pinned(images, out):
            embed_watermark_batch(images, wm, out=out)
    """

    def __init__(self, *arrays):
        self.arrays = [a for a in arrays if a is not None]
        self._done = []

    def __enter__(self):
        lib = _lib.load()
        for a in self.arrays:
            if not (isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]):
                raise ValueError("pinned() takes C-contiguous NumPy arrays")
            _lib.check(lib.tmf_pin_host(a.ctypes.data, a.nbytes))
            self._done.append(a)
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        for a in self._done:
            lib.tmf_unpin_host(a.ctypes.data)
        self._done = []
        return False


_tls = threading.local()


def context_for(device: int, chunk_bytes: int = DEFAULT_CHUNK_BYTES, depth: int = DEFAULT_DEPTH) -> HostPipeline:
    """The calling thread's context for ``device`` (created on first use)."""
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    key = (int(device), int(chunk_bytes), int(depth))
    if key not in cache:
        cache[key] = HostPipeline(*key)
    return cache[key]


def _as_cpu_u8(x, name):
    import torch

    if isinstance(x, np.ndarray):
        if x.dtype != np.uint8:
            raise ValueError(f"{name} must be uint8")
        return torch.from_numpy(np.ascontiguousarray(x)), "numpy"
    if isinstance(x, torch.Tensor) and not x.is_cuda and x.dtype == torch.uint8:
        return x.contiguous(), "torch"
    raise ValueError(f"{name} must be a uint8 NumPy array or CPU torch tensor")


def shard_ranges(n: int, parts: int):
    """Contiguous by-image split: part g gets images [g*n/parts, (g+1)*n/parts)."""
    return [(g * n // parts, (g + 1) * n // parts) for g in range(parts)]


def run_batch(kind, a, b, wm_map, alpha=ALPHA, block_size=BLOCK_SIZE, mode=None,
              devices: Optional[Sequence[int]] = None, out=None, chunk_bytes: int = DEFAULT_CHUNK_BYTES,
              stats: Optional[dict] = None, depth: int = DEFAULT_DEPTH):
    """Batched embed / extract on HOST arrays (NumPy or CPU torch, pinned for full overlap)."""
    from . import watermarking as wmk

    torch = wmk._torch()
    if kind not in ("embed", "extract"):
        raise ValueError("kind must be 'embed' or 'extract'")
    block_size = wmk._require_supported_block(block_size)
    mode = wmk.DEFAULT_MODE if mode is None else int(mode)
    ta, flavour = _as_cpu_u8(a, "images")
    if ta.dim() != 4 or ta.shape[-1] != 3:
        raise ValueError(f"images must have shape (N, H, W, 3), got {tuple(ta.shape)}")
    n, h, w, _ = ta.shape
    nbh, nbw = h // block_size, w // block_size
    tb = None
    if kind == "extract":
        tb, _ = _as_cpu_u8(b, "originals")
        if tb.shape != ta.shape:
            raise ValueError("watermarked and original batches must have the same shape")
    devices = list(devices) if devices is not None else [torch.cuda.current_device()]
    if not devices:
        raise ValueError("devices must not be empty")

    tw, shared = None, True
    if kind == "embed":
        tw, _ = _as_cpu_u8(wm_map, "watermark_map")
        if tuple(tw.shape) == (nbh, nbw):
            shared = True
        elif tuple(tw.shape) == (n, nbh, nbw):
            shared = False
        else:
            raise ValueError(f"watermark map must be {(nbh, nbw)} or {(n, nbh, nbw)}, got {tuple(tw.shape)}")
        out_shape = (n, h, w, 3)
    else:
        out_shape = (n, nbh, nbw)
    if out is None:
        tout = torch.empty(out_shape, dtype=torch.uint8, pin_memory=ta.is_pinned())
    else:
        contiguous = out.flags["C_CONTIGUOUS"] if isinstance(out, np.ndarray) else \
            (isinstance(out, torch.Tensor) and out.is_contiguous())
        if not contiguous:      # a contiguous copy would receive the results and the caller's buffer would stay untouched
            raise ValueError("out must be C-contiguous")
        tout, _ = _as_cpu_u8(out, "out")
        if tuple(tout.shape) != out_shape:
            raise ValueError(f"out must have shape {out_shape}")

    img_bytes, out_bytes, map_bytes = h * w * 3, int(np.prod(out_shape[1:])), nbh * nbw
    # every context is resolved before anything is queued (a bad device id fails here, with nothing in flight)
    plan = [(context_for(dev, chunk_bytes, depth), lo, hi)
            for dev, (lo, hi) in zip(devices, shard_ranges(n, len(devices))) if hi > lo]
    ctxs = []
    try:
        for ctx, lo, hi in plan:
            if stats is not None:
                ctx.stats(reset=True)
            src = ta.data_ptr() + lo * img_bytes
            dst = tout.data_ptr() + lo * out_bytes
            ctxs.append(ctx)
            if kind == "embed":
                wm_ptr = tw.data_ptr() + (0 if shared else lo * map_bytes)
                ctx.embed_async(src, dst, hi - lo, h, w, wm_ptr, shared, alpha, block_size, mode)
            else:
                ctx.extract_async(src, tb.data_ptr() + lo * img_bytes, dst, hi - lo, h, w, alpha, block_size, mode)
    except BaseException:
        # an enqueue failed (out of device memory, ...): the contexts queued so far still copy from / into
        # ta, tb, tw and tout - join them before those buffers can be released, then re-raise
        for ctx in ctxs:
            try:
                ctx.synchronize()
            except Exception:
                pass
        raise
    for ctx in ctxs:       # every device has its work queued before the first wait
        ctx.synchronize()
    if stats is not None:
        tot = {"launches": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        for ctx in ctxs:
            for k, v in ctx.stats().items():
                tot[k] += v
        stats.update(tot)
        stats["chunk_images"] = max(1, min(chunk_bytes // max(1, img_bytes), max(1, -(-n // len(devices)))))
    if out is not None:
        return out
    return tout.numpy() if flavour == "numpy" else tout
