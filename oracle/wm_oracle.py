"""CPU oracle for the DCT+SVD watermark embed/extract path.

TEST INFRASTRUCTURE ONLY.  This module is the checker the CUDA path is compared
against; it is never on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.

It restates, in NumPy, the algorithm of the reference's
``modules/watermarking.py`` (Rigelyon/ThatsMyFace).  Every function cites the
reference lines it follows.  The arithmetic that the reference delegates to
third-party libraries is delegated to the *same* libraries here
(``scipy.fftpack.dct/idct``, ``numpy.linalg.svd`` -> LAPACK sgesdd, ``np.dot``
sgemm, ``PIL.Image.resize(LANCZOS)``; lower bounds in the reference's
``requirements.txt:2-3,7``), so on the same image the block stage is
bit-identical to the reference by construction.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The pin is the reference itself executed in the build
container: ``oracle/make_golden.py`` imports ``/root/reference/modules/watermarking.py``
unmodified (stub ``streamlit``), checks this restatement against it, and writes
``tests/golden/*.npz``; ``tests/test_oracle.py`` re-checks the restatement
against those committed vectors on every run.

Two execution styles are provided:

* ``style="vector"`` (default) - whole-image NumPy; what the parity tests use.
* ``style="loop"`` - the reference's own control flow (one Python iteration per
  pixel for the colour transforms, one per block for DCT/SVD), used by
  ``bench.py`` to time "the reference's CPU path" as the reference performs it.
"""
from __future__ import annotations

import io

import numpy as np
from PIL import Image
from scipy.fftpack import dct, idct

# modules/constants.py:7-8
BLOCK_SIZE = 8
ALPHA = 0.1

# modules/watermarking.py:37-39 and :61 (float64 literals, as in the reference)
RGB2YCC = np.array(
    [[0.299, 0.587, 0.114], [-0.169, -0.331, 0.5], [0.5, -0.419, -0.081]]
)
YCC2RGB = np.array([[1.0, 0.0, 1.403], [1.0, -0.344, -0.714], [1.0, 1.773, 0.0]])


# --------------------------------------------------------------------------
# exact float64 FMA, vectorised (NumPy has none)
# --------------------------------------------------------------------------
_SPLIT = 134217729.0  # 2**27 + 1 (Veltkamp split constant for binary64)


def _two_prod(a, b):
    """Dekker: p + e == a*b exactly (no overflow at these magnitudes)."""
    p = a * b
    ca = _SPLIT * a
    ah = ca - (ca - a)
    al = a - ah
    cb = _SPLIT * b
    bh = cb - (cb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _two_sum(a, b):
    s = a + b
    bb = s - a
    e = (a - (s - bb)) + (b - bb)
    return s, e


def fma64(a, b, c):
    """round_to_nearest_f64(a*b + c) for float64 arrays.

    Error-free product + sum, then one rounding of the three-term tail.  The
    final ``s + (t + e)`` is the correctly rounded result except in
    double-rounding corner cases of probability ~2**-53, irrelevant after the
    cast to float32 that always follows here.
    """
    p, e = _two_prod(np.asarray(a, np.float64), np.asarray(b, np.float64))
    s, t = _two_sum(p, np.asarray(c, np.float64))
    return s + (t + e)


def _dot3_like_npdot(M, v0, v1, v2):
    """Row r of ``np.dot(M, v)`` for a float64 3x3 ``M`` and float64 ``v``.

    ``np.dot`` hands this to OpenBLAS dgemv; measured against the live
    reference in this image (oracle/make_golden.py, exhaustive over all 2**24
    RGB triplets) the accumulation is
    ``fma(M[r,2], v2, fma(M[r,0], v0, M[r,1]*v1))``.
    Returns three float64 arrays (one per output row).
    """
    out = []
    for r in range(3):
        acc = M[r, 1] * v1
        acc = fma64(M[r, 0], v0, acc)
        acc = fma64(M[r, 2], v2, acc)
        out.append(acc)
    return out


# --------------------------------------------------------------------------
# colour transforms
# --------------------------------------------------------------------------
def rgb_to_ycbcr(img, style="vector"):
    """modules/watermarking.py:23-50.

    u8 (or PIL) RGB -> float32 HWC YCbCr: ``x = f32(u8)/255`` (:29), alpha
    channel dropped (:32-34), per pixel ``f32(np.dot(T64, x))`` (:42-45), then
    ``Cb, Cr += 0.5`` in float32 (:48).
    """
    if isinstance(img, Image.Image):
        img = img.convert("RGB")
    x = np.array(img, dtype=np.float32) / 255.0
    if x.shape[-1] == 4:
        x = x[:, :, :3]
    if style == "loop":
        ycc = np.zeros_like(x)
        for i in range(x.shape[0]):
            row = x[i]
            for j in range(x.shape[1]):
                ycc[i, j, :] = np.dot(RGB2YCC, row[j, :])
    else:
        xd = x.astype(np.float64)
        rows = _dot3_like_npdot(RGB2YCC, xd[..., 0], xd[..., 1], xd[..., 2])
        ycc = np.stack(rows, axis=-1).astype(np.float32)
    ycc[:, :, 1:] += 0.5
    return ycc


def ycbcr_to_rgb(ycc, style="vector"):
    """modules/watermarking.py:53-73.

    ``Cb, Cr -= 0.5`` in float32 (:58); per pixel ``f32(np.dot(Ti64, ycc))``
    (:61-67); ``clip(0, 1)`` (:70); ``(rgb * 255).astype(uint8)`` - a float32
    multiply followed by truncation toward zero (:73).
    """
    z = ycc.copy()
    z[:, :, 1:] -= 0.5
    if style == "loop":
        rgb = np.zeros_like(z)
        for i in range(z.shape[0]):
            row = z[i]
            for j in range(z.shape[1]):
                rgb[i, j, :] = np.dot(YCC2RGB, row[j, :])
    else:
        zd = z.astype(np.float64)
        rows = _dot3_like_npdot(YCC2RGB, zd[..., 0], zd[..., 1], zd[..., 2])
        rgb = np.stack(rows, axis=-1).astype(np.float32)
    rgb = np.clip(rgb, 0, 1)
    return (rgb * 255).astype(np.uint8)


# --------------------------------------------------------------------------
# block transforms
# --------------------------------------------------------------------------
def apply_dct_to_block(block):
    """modules/watermarking.py:76-78 - orthonormal 2-D DCT-II, C @ B @ C.T."""
    return dct(dct(block.T, norm="ortho").T, norm="ortho")


def apply_idct_to_block(block):
    """modules/watermarking.py:81-83 - orthonormal 2-D inverse DCT."""
    return idct(idct(block.T, norm="ortho").T, norm="ortho")


def to_blocks(plane, bs=BLOCK_SIZE):
    """(H, W) -> (nbh, nbw, bs, bs) copy of the whole blocks; the partial
    right/bottom strips are left out (``height // block_size``,
    modules/watermarking.py:173-174, :255-256)."""
    h, w = plane.shape
    nbh, nbw = h // bs, w // bs
    v = plane[: nbh * bs, : nbw * bs].reshape(nbh, bs, nbw, bs)
    return np.ascontiguousarray(v.transpose(0, 2, 1, 3))


def from_blocks(blocks, plane):
    """Inverse of ``to_blocks`` - writes the blocks back into ``plane`` in
    place (modules/watermarking.py:207-210)."""
    nbh, nbw, bs, _ = blocks.shape
    plane[: nbh * bs, : nbw * bs] = blocks.transpose(0, 2, 1, 3).reshape(
        nbh * bs, nbw * bs
    )
    return plane


def dct_blocks(blocks):
    """Batched form of :func:`apply_dct_to_block` over the two trailing axes
    (bit-identical to the per-block calls; checked in make_golden.py)."""
    return dct(dct(blocks, axis=-2, norm="ortho"), axis=-1, norm="ortho")


def idct_blocks(blocks):
    return idct(idct(blocks, axis=-2, norm="ortho"), axis=-1, norm="ortho")


def svd_blocks(blocks, vectors=True):
    """``np.linalg.svd(block, full_matrices=True)`` (modules/watermarking.py:195,
    :279-282) batched over leading axes; LAPACK sgesdd for float32 input."""
    if vectors:
        return np.linalg.svd(blocks, full_matrices=True)
    return np.linalg.svd(blocks, full_matrices=True, compute_uv=True)[1]


# --------------------------------------------------------------------------
# watermark preparation (host side in the product too; PIL)
# --------------------------------------------------------------------------
def resize_watermark(watermark, target_height, target_width, preserve_ratio=False):
    """modules/watermarking.py:86-132."""
    wm = Image.open(io.BytesIO(watermark)) if isinstance(watermark, bytes) else watermark
    wm = wm.convert("L")
    if preserve_ratio:
        ow, oh = wm.size
        ratio = min(target_width / ow, target_height / oh)
        nw, nh = int(ow * ratio), int(oh * ratio)
        small = wm.resize((nw, nh), Image.LANCZOS)
        canvas = Image.new("L", (target_width, target_height), 255)
        canvas.paste(small, ((target_width - nw) // 2, (target_height - nh) // 2))
        return canvas
    return wm.resize((target_width, target_height), Image.LANCZOS)


# --------------------------------------------------------------------------
# embed / extract on arrays
# --------------------------------------------------------------------------
def embed_array(rgb_u8, wm_u8, alpha=ALPHA, bs=BLOCK_SIZE, style="vector", taps=None):
    """modules/watermarking.py:163-219 on arrays.

    ``rgb_u8``: (H, W, 3) uint8.  ``wm_u8``: (H//bs, W//bs) uint8, the already
    resized watermark map (:177-180).  Returns (H, W, 3) uint8.
    ``taps``: optional dict that receives intermediates (Y plane, DCT blocks,
    singular values) for stage-level parity tests.
    """
    ycc = rgb_to_ycbcr(rgb_u8, style=style)
    Y = ycc[:, :, 0]
    h, w = Y.shape
    nbh, nbw = h // bs, w // bs
    wm = np.asarray(wm_u8, dtype=np.uint8)
    if wm.shape != (nbh, nbw):
        raise ValueError(f"watermark map must be {(nbh, nbw)}, got {wm.shape}")
    wmf = wm / 255.0  # float64 in [0, 1]  (:180)
    if taps is not None:
        taps["Y"] = Y.copy()

    if style == "loop":
        for i in range(nbh):
            for j in range(nbw):
                blk = Y[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
                d = apply_dct_to_block(blk)
                U, S, Vt = np.linalg.svd(d, full_matrices=True)
                S[0] += alpha * wmf[i, j]
                m = np.dot(U, np.dot(np.diag(S), Vt))
                Y[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = apply_idct_to_block(m)
    elif nbh and nbw:
        B = to_blocks(Y, bs)
        D = dct_blocks(B)
        U, S, Vt = svd_blocks(D)
        if taps is not None:
            taps["D"] = D.copy()
            taps["S"] = S.copy()
            taps["U"] = U.copy()
            taps["Vt"] = Vt.copy()
        # S[0] += alpha * w : float32 + float64 -> float64 -> stored float32 (:198)
        S[..., 0] = (S[..., 0].astype(np.float64) + alpha * wmf).astype(np.float32)
        # np.dot(U, np.dot(np.diag(S), Vt)) per block, with np.dot so that the
        # sgemm summation order is the reference's (:201).  diag(S) @ Vt only
        # adds exact zeros, so it equals the row scaling below bit for bit.
        SV = S[..., :, None] * Vt
        M = np.empty_like(D)
        Uf, SVf, Mf = U.reshape(-1, bs, bs), SV.reshape(-1, bs, bs), M.reshape(-1, bs, bs)
        for k in range(Uf.shape[0]):
            np.dot(Uf[k], SVf[k], out=Mf[k])
        from_blocks(idct_blocks(M), Y)
    ycc[:, :, 0] = Y
    return ycbcr_to_rgb(ycc, style=style)


def sigma0_map(rgb_u8, bs=BLOCK_SIZE, style="vector"):
    """Largest singular value of the DCT of every whole luma block
    (modules/watermarking.py:246-282), float32 (nbh, nbw)."""
    Y = rgb_to_ycbcr(rgb_u8, style=style)[:, :, 0]
    h, w = Y.shape
    nbh, nbw = h // bs, w // bs
    if style == "loop":
        out = np.zeros((nbh, nbw), np.float32)
        for i in range(nbh):
            for j in range(nbw):
                d = apply_dct_to_block(Y[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs])
                out[i, j] = np.linalg.svd(d, full_matrices=True)[1][0]
        return out
    if nbh == 0 or nbw == 0:
        return np.zeros((nbh, nbw), np.float32)
    return svd_blocks(dct_blocks(to_blocks(Y, bs)), vectors=False)[..., 0]


def extract_array(wmk_u8, orig_u8, alpha=ALPHA, bs=BLOCK_SIZE, style="vector"):
    """modules/watermarking.py:242-292 on arrays -> (nbh, nbw) uint8.

    Block grid comes from the *watermarked* image (:254-256).
    ``(S_w[0] - S_o[0]) / alpha`` is float32 under NumPy >= 2 (float32 scalar
    divided by a Python float), then stored into a float64 map (:259, :285);
    ``clip(0, 1)``; ``* 255`` in float64; ``astype(uint8)`` truncates (:288-289).
    """
    h, w = wmk_u8.shape[:2]
    nbh, nbw = h // bs, w // bs
    if orig_u8.shape[:2] != (h, w):
        # the reference slices both images with the watermarked image's block
        # coordinates (:265-272); with unequal sizes that is undefined behaviour
        # (short slices into the DCT), so the oracle only defines equal sizes.
        raise ValueError("watermarked and original images must have the same size")
    sw = sigma0_map(wmk_u8, bs, style)
    so = sigma0_map(orig_u8, bs, style)
    e = ((sw - so) / np.float32(alpha)).astype(np.float64)
    e = np.clip(e, 0, 1)
    return (e * 255).astype(np.uint8)


# --------------------------------------------------------------------------
# PIL-level API with the reference's signatures
# --------------------------------------------------------------------------
def embed_watermark(image, watermark_data, preserve_ratio=False, custom_settings=None, style="vector"):
    """modules/watermarking.py:135-221 (settings: explicit dict or constants;
    the Streamlit session lookup of :10-20 is host-shim logic, not arithmetic)."""
    s = custom_settings or {}
    bs, alpha = s.get("block_size", BLOCK_SIZE), s.get("alpha", ALPHA)
    image = image.convert("RGB")
    wm_img = Image.open(io.BytesIO(watermark_data)) if isinstance(watermark_data, bytes) else watermark_data
    rgb = np.array(image)
    h, w = rgb.shape[:2]
    wm = np.array(resize_watermark(wm_img, h // bs, w // bs, preserve_ratio))
    return Image.fromarray(embed_array(rgb, wm, alpha, bs, style))


def extract_watermark(watermarked_image, original_image, custom_settings=None, style="vector"):
    """modules/watermarking.py:224-294."""
    s = custom_settings or {}
    bs, alpha = s.get("block_size", BLOCK_SIZE), s.get("alpha", ALPHA)
    a = np.array(watermarked_image.convert("RGB"))
    b = np.array(original_image.convert("RGB"))
    return Image.fromarray(extract_array(a, b, alpha, bs, style))
