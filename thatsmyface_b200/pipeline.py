"""Host <-> B200 pipeline around the fused kernels.

The reference's embed page walks the uploaded images one at a time
(``internal_pages/embed_watermark_page.py:492-558``).  Here a batch of host
images is cut into chunks and each chunk flows H2D copy -> fused kernel -> D2H
copy on three CUDA streams with ``depth`` device slots, so PCIe in, compute and
PCIe out overlap.  With several devices the batch is split purely by image
(contiguous ranges, one pipeline per device, no collective, no peer traffic) and
all devices are driven asynchronously from this one host thread.

PyTorch supplies device memory, streams, events and pinned memory only.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .constants import ALPHA, BLOCK_SIZE

DEFAULT_CHUNK_BYTES = 96 << 20


class HostPipeline:
    """Chunked, stream-overlapped embed/extract for one device and one image size."""

    def __init__(self, device: int, h: int, w: int, chunk_images: int, kind: str, depth: int = 2,
                 block_size: int = BLOCK_SIZE):
        from . import watermarking as wmk

        torch = wmk._torch()
        if kind not in ("embed", "extract"):
            raise ValueError("kind must be 'embed' or 'extract'")
        self.torch, self.kind, self.h, self.w = torch, kind, h, w
        self.device = torch.device("cuda", device)
        self.chunk, self.depth = max(1, int(chunk_images)), max(1, int(depth))
        nbh, nbw = h // block_size, w // block_size
        with torch.cuda.device(self.device):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
            mk = lambda *shape: torch.empty(shape, dtype=torch.uint8, device=self.device)
            self.buf_a = [mk(self.chunk, h, w, 3) for _ in range(self.depth)]
            if kind == "embed":
                self.buf_b = None
                self.buf_o = [mk(self.chunk, h, w, 3) for _ in range(self.depth)]
                self.buf_wm = [mk(self.chunk, nbh, nbw) for _ in range(self.depth)]
            else:
                self.buf_b = [mk(self.chunk, h, w, 3) for _ in range(self.depth)]
                self.buf_o = [mk(self.chunk, nbh, nbw) for _ in range(self.depth)]
            self.ev_in = [torch.cuda.Event() for _ in range(self.depth)]
            self.ev_run = [torch.cuda.Event() for _ in range(self.depth)]
            self.ev_free = [torch.cuda.Event() for _ in range(self.depth)]   # slot's input buffers reusable
            self.ev_out = [torch.cuda.Event() for _ in range(self.depth)]    # slot's output buffer reusable
        self.launches = 0
        self._step = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, src_a, src_b, dst, wm, alpha, block_size, mode):
        """Queue one chunk (<= self.chunk images).  ``src_*``/``dst`` are CPU uint8
        tensors (pinned for true asynchrony).  ``wm`` (embed only) is either a
        device-resident shared map ``(nbh, nbw)`` or a CPU tensor of per-image maps
        ``(k, nbh, nbw)``, which then rides the H2D stream with the images.
        Returns immediately."""
        from . import watermarking as wmk

        torch = self.torch
        k = src_a.shape[0]
        slot = self._step % self.depth
        first_use = self._step < self.depth
        self._step += 1
        with torch.cuda.device(self.device):
            a = self.buf_a[slot][:k]
            o = self.buf_o[slot][:k]
            with torch.cuda.stream(self.s_in):
                if not first_use:
                    self.s_in.wait_event(self.ev_free[slot])
                a.copy_(src_a, non_blocking=True)
                self.h2d_bytes += src_a.numel()
                if self.kind == "extract":
                    b = self.buf_b[slot][:k]
                    b.copy_(src_b, non_blocking=True)
                    self.h2d_bytes += src_b.numel()
                wm_dev = wm
                if self.kind == "embed" and not wm.is_cuda:
                    wm_dev = self.buf_wm[slot][:k]
                    wm_dev.copy_(wm, non_blocking=True)
                    self.h2d_bytes += wm.numel()
                self.ev_in[slot].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[slot])
                if not first_use:
                    self.s_run.wait_event(self.ev_out[slot])
                if self.kind == "embed":
                    wmk.embed_tensor(a, wm_dev, alpha, block_size, mode, out=o)
                else:
                    wmk.extract_tensor(a, self.buf_b[slot][:k], alpha, block_size, mode, out=o)
                self.launches += 1
                self.ev_run[slot].record(self.s_run)
                self.ev_free[slot].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[slot])
                dst.copy_(o, non_blocking=True)
                self.d2h_bytes += o.numel()
                self.ev_out[slot].record(self.s_out)

    def synchronize(self):
        self.s_out.synchronize()


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
              stats: Optional[dict] = None):
    from . import watermarking as wmk

    torch = wmk._torch()
    block_size = wmk._require_supported_block(block_size)
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

    if kind == "embed":
        tw, _ = _as_cpu_u8(wm_map, "watermark_map")
        if tuple(tw.shape) not in ((nbh, nbw), (n, nbh, nbw)):
            raise ValueError(f"watermark map must be {(nbh, nbw)} or {(n, nbh, nbw)}, got {tuple(tw.shape)}")
        out_shape = (n, h, w, 3)
    else:
        tw = None
        out_shape = (n, nbh, nbw)
    if out is None:
        tout = torch.empty(out_shape, dtype=torch.uint8, pin_memory=ta.is_pinned())
    else:
        tout, _ = _as_cpu_u8(out, "out")
        if tuple(tout.shape) != out_shape:
            raise ValueError(f"out must have shape {out_shape}")

    img_bytes = max(1, h * w * 3)
    chunk = max(1, min(chunk_bytes // img_bytes, max(1, -(-n // len(devices)))))
    pipes, wms, work = [], [], []
    for dev, (lo, hi) in zip(devices, shard_ranges(n, len(devices))):
        if hi <= lo:
            continue
        p = HostPipeline(dev, h, w, min(chunk, hi - lo), kind, block_size=block_size)
        pipes.append(p)
        if kind == "embed" and tw.dim() == 2:
            wms.append(tw.to(p.device))          # shared map: resident before the first launch
            torch.cuda.current_stream(p.device).synchronize()
        else:
            wms.append(None)
        work.append([(s, min(s + p.chunk, hi)) for s in range(lo, hi, p.chunk)])
    # round-robin over devices so that every device has work queued early
    for step in range(max((len(wk) for wk in work), default=0)):
        for p, wm_dev, wk in zip(pipes, wms, work):
            if step >= len(wk):
                continue
            s, e = wk[step]
            wm_chunk = tw[s:e] if (kind == "embed" and wm_dev is None) else wm_dev
            p.submit(ta[s:e], tb[s:e] if tb is not None else None, tout[s:e], wm_chunk, alpha, block_size, mode)
    for p in pipes:
        p.synchronize()
    if stats is not None:
        stats["launches"] = sum(p.launches for p in pipes)
        stats["h2d_bytes"] = sum(p.h2d_bytes for p in pipes)
        stats["d2h_bytes"] = sum(p.d2h_bytes for p in pipes)
        stats["chunk_images"] = chunk
    if out is not None:
        return out
    return tout.numpy() if flavour == "numpy" else tout
