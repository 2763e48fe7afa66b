"""CPU restatement of Pillow's 8-bit LANCZOS resampler (mode "L").

TEST INFRASTRUCTURE ONLY (same rule as ``wm_oracle.py``): the checker for
``tmf_wm_map_l8``, never on the product path.

The reference resizes the watermark with ``PIL.Image.resize(size, Image.LANCZOS)``
on a mode-"L" image (``modules/watermarking.py:113`` and ``:128-130``).  Pillow is a
third-party dependency that is not under ``/root/reference`` (``requirements.txt:3``
``Pillow>=9.0.0``, unpinned; 12.2.0 installed here), so the algorithm is restated from
its published C source, ``src/libImaging/Resample.c``:

* ``precompute_coeffs``: per output sample ``xx`` the window
  ``[xmin, xmin+xmax)`` = ``[int(c - s + .5), int(c + s + .5))`` clipped to the
  input, with ``c = (xx + .5)*scale``, ``s = 3*max(scale, 1)``; float64 weights
  ``lanczos(((x + xmin) - c + .5) / max(scale, 1))`` normalised by their sum;
* ``normalize_coeffs_8bpc``: weights to fixed point, ``int(+-0.5 + w * 2**22)``;
* ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``:
  ``out = clip8((2**21 + sum(pixel * k)) >> 22)`` in int32, the horizontal pass first,
  rounding to uint8 between the passes, only over the source rows the vertical pass
  needs;
* ``Image.resize`` (Python): identical size returns a copy; images more than 100x
  taller than wide resample vertically first.

Parity pin: ``tests/test_wm_map.py`` checks this file against the installed Pillow
itself on fixed and hypothesis-drawn sizes (bit-exact), and the CUDA kernel against
this file.  ``math.sin`` is the C library's ``sin``, the same one Pillow's extension
calls, so the float64 weights and therefore the fixed-point tables are identical.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x: float) -> float:
    """Resample.c lanczos_filter: truncated sinc(x)*sinc(x/3) on [-3, 3)."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).

    Returns (ksize, bounds[out_size, 2] = (xmin, count), kk[out_size, ksize] int32).
    """
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis1(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass along the last axis (the horizontal pass; the vertical one on the transpose)."""
    _, bounds, kk = precompute_coeffs(img.shape[1], out_size)
    out = np.empty((img.shape[0], out_size), np.uint8)
    src = img.astype(np.int32)
    for xx in range(out_size):
        x0, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + src[:, x0:x0 + cnt] @ kk[xx, :cnt]
        out[:, xx] = _clip8(acc.astype(np.int32))
    return out


def resize_l8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``Image.fromarray(img, "L").resize((out_w, out_h), Image.LANCZOS)`` as an array."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim != 2:
        raise ValueError("mode L image expected")
    if out_h <= 0 or out_w <= 0:
        raise ValueError("height and width must be > 0")
    in_h, in_w = img.shape
    if (in_h, in_w) == (out_h, out_w):
        return img.copy()
    if in_h > in_w * 100 and out_h < in_h:   # Image.resize: very tall images go vertical first
        tmp = _resample_axis1(img.T, out_h).T if out_h != in_h else img
        return _resample_axis1(tmp, out_w) if out_w != in_w else tmp.copy()
    cur = img
    if out_w != in_w:
        # Resample.c ImagingResampleInner: the horizontal pass only covers the source rows the
        # vertical pass reads; rows are independent, so resampling all of them gives the same pixels
        cur = _resample_axis1(cur, out_w)
    if out_h != in_h:
        cur = _resample_axis1(np.ascontiguousarray(cur.T), out_h).T
    return np.ascontiguousarray(cur)


def watermark_geometry(src_h: int, src_w: int, target_h: int, target_w: int, preserve_ratio: bool):
    """Sizes and paste offset of resize_watermark (modules/watermarking.py:105-123).

    Returns (new_h, new_w, paste_y, paste_x)."""
    if preserve_ratio:
        ratio = min(target_w / src_w, target_h / src_h)
        new_w, new_h = int(src_w * ratio), int(src_h * ratio)
        return new_h, new_w, (target_h - new_h) // 2, (target_w - new_w) // 2
    return target_h, target_w, 0, 0


def watermark_map_l8(img: np.ndarray, target_h: int, target_w: int, preserve_ratio: bool = False) -> np.ndarray:
    """resize_watermark (modules/watermarking.py:86-132) after ``.convert("L")``, on arrays."""
    new_h, new_w, py, px = watermark_geometry(img.shape[0], img.shape[1], target_h, target_w, preserve_ratio)
    small = resize_l8(img, new_h, new_w)
    if not preserve_ratio:
        return small
    canvas = np.full((target_h, target_w), 255, np.uint8)
    canvas[py:py + new_h, px:px + new_w] = small     # Image.paste at (paste_x, paste_y); always inside
    return canvas
