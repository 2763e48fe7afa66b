"""The pages' per-image loops as one call (the caller side).

The reference's embed page opens each upload, calls ``embed_watermark(img, watermark_data,
preserve_ratio=True)`` and PNG-encodes the result, strictly one image after the other
(``internal_pages/embed_watermark_page.py:492-558``; ``:533-534`` is the PNG encode); the extract
page does the same with ``extract_watermark`` (``extract_watermark_page.py:293-296``).  Once the
block loop runs on the B200 in tens of microseconds, what is left per image is host work: file
decode, ``convert("RGB")``, the PCIe copies and - by far the largest - the PNG encode.

``embed_watermark_many`` / ``extract_watermark_many`` keep the per-image semantics (images of ANY
size and mode, same results as the single-image functions, input order preserved) and run the
images on a small pool of *lanes*.  A lane is one host thread with its own CUDA stream and its own
pinned staging buffers: decode -> stage -> H2D -> fused kernel -> D2H -> PIL / PNG.  PIL's codecs,
NumPy's copies, the ctypes call and the stream wait all release the GIL, so the lanes overlap
each other's host work with the copies and kernels, and PNG encoding leaves the critical path.
With several devices the lanes are dealt round-robin over them (by-image sharding, no collective).

There is no CPU fallback: without a CUDA device the calls raise ``RuntimeError``.
"""
from __future__ import annotations

import concurrent.futures
import io
import os
import threading
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np
from PIL import Image

from . import watermarking as wmk

__all__ = ["embed_watermark_many", "embed_watermark_iter", "extract_watermark_many", "extract_watermark_iter"]

DEFAULT_LANES = max(2, min(16, os.cpu_count() or 2))   # near-linear to 16 lanes on a 16-core host (profiles/r01_page_loop.jsonl)


class _Lane:
    """Per-thread resources: device, stream, two pinned staging buffers that only grow."""

    def __init__(self, torch, device: int):
        self.torch = torch
        self.device = int(device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream()
        self._pinned = {}

    def staging(self, name: str, nbytes: int):
        buf = self._pinned.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = self.torch.empty(max(nbytes, 1 << 20), dtype=self.torch.uint8, pin_memory=True)
            self._pinned[name] = buf
        return buf[:nbytes]

    def upload(self, name: str, arr: np.ndarray):
        """host array -> pinned staging -> device tensor on this lane's stream"""
        flat = self.staging(name, arr.size)
        np.copyto(flat.numpy(), arr.reshape(-1))
        return flat.view(arr.shape).cuda(non_blocking=True)

    def download(self, name: str, t) -> np.ndarray:
        """device tensor -> pinned staging -> fresh host array (the staging buffer is reused)"""
        flat = self.staging(name, t.numel())
        flat.view(t.shape).copy_(t, non_blocking=True)
        self.stream.synchronize()
        return flat.view(t.shape).numpy().copy()


class _Lanes:
    def __init__(self, devices: Optional[Sequence[int]]):
        self.torch = wmk._torch()
        self.devices = [int(d) for d in devices] if devices else [self.torch.cuda.current_device()]
        if not self.devices:
            raise ValueError("devices must not be empty")
        self._tls = threading.local()
        self._count = 0
        self._lock = threading.Lock()

    def mine(self) -> _Lane:
        lane = getattr(self._tls, "lane", None)
        if lane is None:
            with self._lock:
                k = self._count
                self._count += 1
            lane = self._tls.lane = _Lane(self.torch, self.devices[k % len(self.devices)])
        return lane


def _open(item) -> Image.Image:
    if isinstance(item, Image.Image):
        return item
    if isinstance(item, (bytes, bytearray)):
        return Image.open(io.BytesIO(bytes(item)))
    return Image.open(item)          # path or file-like (what the page passes, :518)


def _ordered(pool_size: int, fn, items) -> Iterator:
    """fn over items on a thread pool, results in input order, at most 2*pool_size in flight."""
    items = iter(items)
    with concurrent.futures.ThreadPoolExecutor(pool_size, thread_name_prefix="tmf-lane") as pool:
        pending = []
        try:
            for item in items:
                pending.append(pool.submit(fn, item))
                if len(pending) >= 2 * pool_size:
                    yield pending.pop(0).result()
            while pending:
                yield pending.pop(0).result()
        finally:
            for f in pending:
                f.cancel()


def embed_watermark_iter(images: Iterable, watermark_data, preserve_ratio: bool = False, custom_settings=None, *,
                         png: bool = False, lanes: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                         png_options: Optional[dict] = None) -> Iterator:
    """Generator form of ``embed_watermark_many`` (results arrive in input order as they finish,
    e.g. to drive the page's progress bar)."""
    block_size, alpha, mode = wmk._resolve(custom_settings)
    wmk._require_supported_block(block_size)
    pool = _Lanes(devices)
    torch = pool.torch
    opts = dict(png_options or {})

    def one(item):
        lane = pool.mine()
        image = _open(item)
        if wmk._fast_pil():
            # PIL's 4-byte pixels cross PCIe as they are (watermarking.py, "PIL boundary"): the lane's
            # thread owns its staging blocks, the conversions run on the lane's stream
            with torch.cuda.device(lane.device), torch.cuda.stream(lane.stream):
                x = wmk._pil_to_device_rgb(image)
                h, w = x.shape[:2]
                m = wmk.watermark_map(watermark_data, h // block_size, w // block_size, preserve_ratio,
                                      device=torch.device("cuda", lane.device))
                result = wmk._device_rgb_to_pil(wmk.embed_tensor(x, m, alpha, block_size, mode))
        else:
            rgb = wmk._pil_to_rgb_array(image)
            h, w = rgb.shape[:2]
            with torch.cuda.device(lane.device), torch.cuda.stream(lane.stream):
                x = lane.upload("in", rgb)
                m = wmk.watermark_map(watermark_data, h // block_size, w // block_size, preserve_ratio,
                                      device=torch.device("cuda", lane.device))
                out = lane.download("out", wmk.embed_tensor(x, m, alpha, block_size, mode))
            result = wmk._array_to_pil(out, "RGB")
        if not png:
            return result
        buf = io.BytesIO()
        result.save(buf, format="PNG", **opts)          # embed_watermark_page.py:533-534
        return result, buf.getvalue()

    return _ordered(int(lanes or DEFAULT_LANES), one, images)


def embed_watermark_many(images: Iterable, watermark_data, preserve_ratio: bool = False, custom_settings=None, *,
                         png: bool = False, lanes: Optional[int] = None, devices: Optional[Sequence[int]] = None,
                         png_options: Optional[dict] = None) -> list:
    """``[embed_watermark(img, watermark_data, preserve_ratio, custom_settings) for img in images]``
    with the images in flight on ``lanes`` host threads / CUDA streams.

    ``images``: PIL images, paths, file objects or encoded ``bytes`` (anything the page hands to
    ``Image.open``), of any sizes and modes.  Returns a list of PIL "RGB" images, or with
    ``png=True`` a list of ``(image, png_bytes)`` - the two things the page keeps per upload
    (``embed_watermark_page.py:531-545``); ``png_options`` go to ``Image.save``."""
    return list(embed_watermark_iter(images, watermark_data, preserve_ratio, custom_settings, png=png, lanes=lanes,
                                     devices=devices, png_options=png_options))


def extract_watermark_iter(pairs: Iterable[Tuple], custom_settings=None, *, lanes: Optional[int] = None,
                           devices: Optional[Sequence[int]] = None) -> Iterator:
    block_size, alpha, mode = wmk._resolve(custom_settings)
    wmk._require_supported_block(block_size)
    pool = _Lanes(devices)
    torch = pool.torch

    def one(pair):
        lane = pool.mine()
        if wmk._fast_pil():
            ia, ib = _open(pair[0]), _open(pair[1])
            if ia.size != ib.size:
                raise ValueError(f"watermarked image {ia.size[0]}x{ia.size[1]} and original image "
                                 f"{ib.size[0]}x{ib.size[1]} must have the same size")
            with torch.cuda.device(lane.device), torch.cuda.stream(lane.stream):
                out = lane.download("out", wmk.extract_tensor(wmk._pil_to_device_rgb(ia), wmk._pil_to_device_rgb(ib),
                                                               alpha, block_size, mode))
            return wmk._array_to_pil(out, "L")
        a = wmk._pil_to_rgb_array(_open(pair[0]))
        b = wmk._pil_to_rgb_array(_open(pair[1]))
        if a.shape != b.shape:
            raise ValueError(f"watermarked image {a.shape[1]}x{a.shape[0]} and original image "
                             f"{b.shape[1]}x{b.shape[0]} must have the same size")
        with torch.cuda.device(lane.device), torch.cuda.stream(lane.stream):
            out = lane.download("out", wmk.extract_tensor(lane.upload("in", a), lane.upload("in2", b), alpha,
                                                           block_size, mode))
        return wmk._array_to_pil(out, "L")

    return _ordered(int(lanes or DEFAULT_LANES), one, pairs)


def extract_watermark_many(pairs: Iterable[Tuple], custom_settings=None, *, lanes: Optional[int] = None,
                           devices: Optional[Sequence[int]] = None) -> list:
    """``[extract_watermark(w, o, custom_settings) for (w, o) in pairs]`` on the lane pool; each
    result is a PIL "L" image of size ``(W//bs, H//bs)`` (extract_watermark_page.py:293-296)."""
    return list(extract_watermark_iter(pairs, custom_settings, lanes=lanes, devices=devices))
