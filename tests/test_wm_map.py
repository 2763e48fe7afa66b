"""Watermark-map row: resize_watermark after ``.convert("L")``.

The bar is bit-exactness (integer work).  Three layers, each against the layer above it:

  installed Pillow / committed maps of the live reference (tests/golden/wm_map_cases.npz)
    <- oracle/pil_lanczos.py (NumPy restatement of Pillow's Resample.c)
    <- the library's own table builder + per-sample function run on the host (tests/hostsim)
    <- tmf_wm_map_l8 on the B200 through the C ABI (gpu-marked).
"""
import io
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from PIL import Image

import hostsim_util as hs
from conftest import GOLDEN_DIR
from oracle import pil_lanczos as PL
from oracle import wm_oracle as O


def golden_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "wm_map_cases.npz"))
    meta = json.load(open(os.path.join(GOLDEN_DIR, "MANIFEST.json")))["wm_map_cases"]["cases"]
    return [(m, z["src_" + m["source"]], z[f"map_{i}"]) for i, m in enumerate(meta)]


GOLDEN = golden_cases()
IDS = [f'{m["source"]}-{m["target_h"]}x{m["target_w"]}-pr{int(m["preserve_ratio"])}' for m, _, _ in GOLDEN]


def pil_map(src, th, tw, pr):
    """The reference's function body (via the oracle's restatement of it, which calls PIL itself)."""
    return np.array(O.resize_watermark(Image.fromarray(src, "L"), th, tw, pr))


def source(kind, h, w, seed):
    rng = np.random.default_rng(seed)
    if kind == 0:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 1:   # QR-like: 7-pixel black/white modules
        cells = (rng.integers(0, 2, ((h + 6) // 7, (w + 6) // 7)) * 255).astype(np.uint8)
        return np.ascontiguousarray(np.kron(cells, np.ones((7, 7), np.uint8))[:h, :w])
    return np.full((h, w), int(rng.integers(0, 256)), np.uint8)


# ----------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", GOLDEN, ids=IDS)
def test_restatement_matches_the_live_reference_maps(case):
    m, src, ref = case
    assert np.array_equal(PL.watermark_map_l8(src, m["target_h"], m["target_w"], m["preserve_ratio"]), ref)


@pytest.mark.parametrize("case", GOLDEN, ids=IDS)
def test_library_host_arithmetic_matches_the_live_reference_maps(case):
    m, src, ref = case
    assert np.array_equal(hs.wm_map_l8(src, m["target_h"], m["target_w"], m["preserve_ratio"]), ref)


def test_png_goldens_of_the_embed_cases(golden):
    """pil_png_preserve{0,1}.npz hold the PNG bytes and the map the live reference made of them."""
    for name, pr in (("pil_png_preserve0", False), ("pil_png_preserve1", True)):
        g = golden(name)
        src = np.array(Image.open(io.BytesIO(g["png"].tobytes())).convert("L"))
        th, tw = g["wm"].shape
        assert np.array_equal(PL.watermark_map_l8(src, th, tw, pr), g["wm"])
        assert np.array_equal(hs.wm_map_l8(src, th, tw, pr), g["wm"])


@settings(max_examples=120, deadline=None)
@given(st.integers(1, 260), st.integers(1, 260), st.integers(1, 200), st.integers(1, 200), st.booleans(),
       st.integers(0, 2), st.integers(0, 2**31 - 1))
def test_restatement_and_library_table_builder_match_installed_pillow(sh, sw, th, tw, pr, kind, seed):
    src = source(kind, sh, sw, seed)
    try:
        ref = pil_map(src, th, tw, pr)
    except ValueError:           # a side of the ratio-preserving size truncated to 0: PIL raises
        with pytest.raises(ValueError):
            PL.watermark_map_l8(src, th, tw, pr)
        with pytest.raises(ValueError):
            hs.wm_map_l8(src, th, tw, pr)
        return
    assert np.array_equal(PL.watermark_map_l8(src, th, tw, pr), ref)
    if not (sh > 100 * sw):      # the library refuses the vertical-first case
        assert np.array_equal(hs.wm_map_l8(src, th, tw, pr), ref)


def test_weight_tables_are_what_pillow_uses():
    """Coefficients sum to 2**22 up to rounding and windows are clipped to the image."""
    for in_size, out_size in ((1000, 135), (135, 1000), (64, 64), (7, 3), (3, 7)):
        ksize, bounds, kk = PL.precompute_coeffs(in_size, out_size)
        assert kk.shape == (out_size, ksize)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= in_size).all()
        assert np.abs(kk.sum(axis=1) - (1 << PL.PRECISION_BITS)).max() <= ksize


def test_tall_sources_take_pillows_vertical_first_path():
    src = source(0, 1300, 4, 3)
    ref = np.array(Image.fromarray(src, "L").resize((3, 40), Image.LANCZOS))
    assert np.array_equal(PL.resize_l8(src, 40, 3), ref)


def product_axis_table(in_size, out_size):
    """The PRODUCT library's own host table (tmf_wm_map_axis_table: host only, no device touched)."""
    import ctypes as C

    from thatsmyface_b200 import _lib
    lib = _lib.load()
    ks = C.c_int(0)
    assert lib.tmf_wm_map_axis_table(in_size, out_size, C.byref(ks), None, None, 0) == 0
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ks.value), np.int32)
    assert lib.tmf_wm_map_axis_table(in_size, out_size, C.byref(ks), bounds.ctypes.data, kk.ctypes.data, kk.size) == 0
    return ks.value, bounds, kk


@pytest.mark.parametrize("in_size,out_size", [(290, 240), (290, 135), (370, 64), (1000, 7), (33, 240), (8, 8), (49136, 5),
                                              (777, 270), (1, 3), (3, 1), (4096, 480)])
def test_product_library_host_tables_equal_pillows(in_size, out_size):
    """libtmfwm.so's host floating point (libm sin, float64 weights rounded to 22-bit fixed point,
    built with -ffp-contract=off) must give Pillow's tables bit for bit - checked on the shipped .so, not on
    the hostsim build.  The restatement it is compared with is pinned to Pillow above; the second half asks
    Pillow directly: a one-row image through the tables equals Image.resize."""
    ks, bounds, kk = product_axis_table(in_size, out_size)
    ks_o, bounds_o, kk_o = PL.precompute_coeffs(in_size, out_size)
    assert ks == ks_o
    assert np.array_equal(bounds, bounds_o)
    assert np.array_equal(kk, kk_o)
    rng = np.random.default_rng(in_size * 31 + out_size)
    row = rng.integers(0, 256, (1, in_size), dtype=np.uint8)
    acc = np.array([(row[0, b0:b0 + n].astype(np.int64) * kk[i, :n]).sum() for i, (b0, n) in enumerate(bounds)])
    mine = np.clip((acc + (1 << (PL.PRECISION_BITS - 1))) >> PL.PRECISION_BITS, 0, 255).astype(np.uint8)
    pil = np.array(Image.fromarray(row, "L").resize((out_size, 1), Image.LANCZOS))[0]
    assert np.array_equal(mine, pil)


def test_axis_table_rejects_bad_sizes():
    import ctypes as C

    from thatsmyface_b200 import _lib
    lib = _lib.load()
    ks = C.c_int(0)
    assert lib.tmf_wm_map_axis_table(0, 4, C.byref(ks), None, None, 0) == _lib.ERR_BAD_ARG
    small = np.zeros(4, np.int32)
    assert lib.tmf_wm_map_axis_table(100, 10, C.byref(ks), small.ctypes.data, small.ctypes.data, 4) == _lib.ERR_BAD_ARG


# ----------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")


def gpu_map(src, th, tw, pr):
    from thatsmyface_b200 import watermarking as W

    t = torch.from_numpy(np.ascontiguousarray(src)).cuda()
    return W.watermark_map_tensor(t, th, tw, pr).cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN, ids=IDS)
def test_gpu_matches_the_live_reference_maps(case):
    m, src, ref = case
    assert np.array_equal(gpu_map(src, m["target_h"], m["target_w"], m["preserve_ratio"]), ref)


@pytest.mark.gpu
def test_gpu_matches_installed_pillow_on_random_shapes():
    rng = np.random.default_rng(5)
    checked = 0
    for t in range(150):
        sh, sw, th, tw = (int(v) for v in rng.integers(1, 420, 4))
        pr = bool(rng.integers(0, 2))
        src = source(int(rng.integers(0, 3)), sh, sw, t)
        if sh > 100 * sw:
            continue
        try:
            ref = pil_map(src, th, tw, pr)
        except ValueError:
            with pytest.raises(ValueError):
                gpu_map(src, th, tw, pr)
            continue
        got = gpu_map(src, th, tw, pr)
        assert np.array_equal(got, ref), (sh, sw, th, tw, pr)
        checked += 1
    assert checked > 100


@pytest.mark.gpu
@pytest.mark.parametrize("shape,target,pr", [
    ((1000, 1000), (135, 240), True),      # the page's QR on a 1080p image
    ((1000, 1000), (270, 480), True),      # ... on a 4K image
    ((1000, 1000), (1000, 1000), False),   # same size: PIL returns a copy
    ((1000, 999), (1000, 240), False),     # horizontal pass only, odd width (unaligned rows)
    ((333, 240), (135, 240), False),       # vertical pass only
    ((50, 70), (135, 240), True),          # upscale
    ((40, 7001), (20, 300), False),        # 4 rows per CTA
    ((12, 20011), (12, 500), False),       # 2 rows per CTA
    ((5, 30001), (3, 700), False),         # 1 row per CTA
])
def test_gpu_pass_selection_and_wide_sources(shape, target, pr):
    src = source(1, shape[0], shape[1], 9)
    assert np.array_equal(gpu_map(src, target[0], target[1], pr), pil_map(src, target[0], target[1], pr))


@pytest.mark.gpu
def test_gpu_batch_of_distinct_watermarks_and_out_argument():
    from thatsmyface_b200 import watermarking as W

    srcs = np.stack([source(k % 3, 211, 190, k) for k in range(7)])
    ref = np.stack([pil_map(s, 135, 240, True) for s in srcs])
    t = torch.from_numpy(srcs).cuda()
    got = W.watermark_map_tensor(t, 135, 240, True)
    assert got.shape == (7, 135, 240) and np.array_equal(got.cpu().numpy(), ref)
    out = torch.zeros((7, 135, 240), dtype=torch.uint8, device="cuda")
    assert W.watermark_map_tensor(t, 135, 240, True, out=out) is out
    assert np.array_equal(out.cpu().numpy(), ref)
    # an unaligned view of the batch (image 1 onwards starts at an odd address)
    sub = t.flatten()[211 * 190:].view(6, 211, 190)
    assert np.array_equal(W.watermark_map_tensor(sub, 135, 240, True).cpu().numpy(), ref[1:])


@pytest.mark.gpu
def test_gpu_watermark_maps_from_png_bytes_of_mixed_sizes_feed_embed():
    from thatsmyface_b200 import watermarking as W

    pngs = []
    for k, (h, w) in enumerate([(300, 300), (123, 77), (300, 300), (64, 200), (123, 77)]):
        buf = io.BytesIO()
        mode_img = Image.fromarray(source(1, h, w, 20 + k), "L")
        (mode_img.convert("RGB") if k % 2 else mode_img).save(buf, format="PNG")
        pngs.append(buf.getvalue())
    maps = W.watermark_maps(pngs, 16, 25, preserve_ratio=True)
    ref = np.stack([np.array(W.resize_watermark(p, 16, 25, True)) for p in pngs])
    assert np.array_equal(maps.cpu().numpy(), ref)
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (5, 128, 200, 3), dtype=np.uint8)
    x = torch.from_numpy(rgb).cuda()
    a = W.embed_tensor(x, maps)
    b = W.embed_tensor(x, torch.from_numpy(ref).cuda())
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_gpu_errors_follow_pillow_and_the_header():
    from thatsmyface_b200 import _lib
    from thatsmyface_b200 import watermarking as W

    t = torch.zeros((1, 10, 1000), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError, match="must be > 0"):
        W.watermark_map_tensor(t, 5, 50, True)          # ratio 0.05 -> height int(0.5) = 0
    with pytest.raises(ValueError, match="100x taller"):
        W.watermark_map_tensor(torch.zeros((1, 900, 2), dtype=torch.uint8, device="cuda"), 40, 2, False)
    with pytest.raises(ValueError):
        W.watermark_map_tensor(torch.zeros((10, 10), dtype=torch.float32, device="cuda"), 5, 5)
    lib = _lib.load()
    need = lib.tmf_wm_map_workspace_bytes(1, 10, 1000, 5, 500, 0)
    assert need > 0
    maps = torch.empty((1, 5, 500), dtype=torch.uint8, device="cuda")
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    rc = lib.tmf_wm_map_l8(t.data_ptr(), 1, 10, 1000, 10000, maps.data_ptr(), 5, 500, 0, small.data_ptr(), 16, None)
    assert rc == -1 and "workspace" in _lib.last_error()
    assert lib.tmf_wm_map_l8(t.data_ptr(), 0, 10, 1000, 10000, maps.data_ptr(), 5, 500, 0, None, 0, None) == 0
