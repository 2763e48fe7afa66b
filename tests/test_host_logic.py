"""CPU-side checks: the C-ABI library loads and exports what include/tmf_wm.h
declares, the host shim mirrors the reference's settings / resize / error
behaviour, and nothing computes without a GPU (no CPU fallback)."""
import io
import os
import re
import sys
import types

import numpy as np
import pytest
from PIL import Image

from conftest import ROOT
from oracle import wm_oracle as O
from thatsmyface_b200 import _lib, build as tmf_build
from thatsmyface_b200 import watermarking as W
from thatsmyface_b200.pipeline import shard_ranges

torch = pytest.importorskip("torch")
NO_GPU = not torch.cuda.is_available()


@pytest.fixture(scope="module")
def lib():
    tmf_build.build()          # no-op when the in-tree .so is current
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "tmf_wm.h")).read()
    declared = set(re.findall(r"\b(tmf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.tmf_version() == 201


def test_argument_validation_happens_before_any_cuda_work(lib):
    # these return from the host-side checks, so they are safe without a GPU
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 16 * 16 * 3, None, 1, 0.1, 5, 0, None) == -2
    assert "block_size 5 is not supported" in _lib.last_error()
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 16 * 16 * 3, None, 1, 0.1, 18, 0, None) == -2
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 16 * 16 * 3, None, 1, 0.1, 4, 0, None) == -1   # supported size, null pointers
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 16 * 16 * 3, None, 1, 0.1, 8, 0, None) == -1
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 10, None, 1, 0.1, 8, 0, None) == -1
    assert "img_stride" in _lib.last_error()
    assert lib.tmf_embed_rgb8(None, None, 1, 16, 16, 768, None, 1, 0.1, 8, 7, None) == -1
    assert lib.tmf_extract_rgb8(None, None, None, 1, 16, 16, 768, 0.0, 8, 0, None) == -1
    assert lib.tmf_svd8x8_f32(None, -1, None, None, None, None, 0, None) == -1
    assert lib.tmf_svd8x8_f32(None, 0, None, None, None, None, 0, None) == 0
    assert lib.tmf_embed_rgb8(None, None, 0, 16, 16, 768, None, 1, 0.1, 8, 0, None) == 0   # empty batch
    import ctypes as C
    h = C.c_void_p()
    assert lib.tmf_ctx_create(C.byref(h), 0, 0, 99) == -1 and "depth" in _lib.last_error()
    assert lib.tmf_ctx_create(None, 0, 0, 2) == -1
    if NO_GPU:
        assert lib.tmf_ctx_create(C.byref(h), 0, 0, 2) == -3 and not h      # no device: CUDA error, no context
    assert lib.tmf_ctx_embed_host_async(None, None, None, 1, 16, 16, None, 1, 0.1, 8, 1) == -1
    assert lib.tmf_ctx_synchronize(None) == -1 and lib.tmf_ctx_destroy(None) == 0
    assert lib.tmf_pin_host(None, 0) == -1
    with pytest.raises(ValueError):
        _lib.check(-1)
    with pytest.raises(RuntimeError):
        _lib.check(-3)


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    img = Image.new("RGB", (16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.embed_watermark(img, Image.new("L", (2, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.extract_watermark(img, img)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.rgb_to_ycbcr(img)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        W.embed_watermark_batch(np.zeros((1, 16, 16, 3), np.uint8), np.zeros((2, 2), np.uint8))


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for top in ("thatsmyface_b200", "include", "profiles"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in src.replace("oracle/make_golden.py", ""), os.path.join(dirpath, f)


def test_settings_precedence(monkeypatch):
    """explicit truthy dict -> st.session_state.custom_settings -> constants
    (modules/watermarking.py:10-20, :149-151)."""
    monkeypatch.delitem(sys.modules, "streamlit", raising=False)
    assert W.get_watermark_settings() == {"block_size": 8, "alpha": 0.1}
    assert W._resolve(None)[:2] == (8, 0.1)
    assert W._resolve({})[:2] == (8, 0.1)                      # falsy dict falls through
    assert W._resolve({"alpha": 0.5})[:2] == (8, 0.5)
    st = types.ModuleType("streamlit")

    class State(dict):
        __getattr__ = dict.__getitem__

    st.session_state = State()
    monkeypatch.setitem(sys.modules, "streamlit", st)
    assert W.get_watermark_settings() == {"block_size": 8, "alpha": 0.1}     # key absent
    st.session_state["custom_settings"] = {"alpha": 0.3}
    assert W.get_watermark_settings() == {"block_size": 8, "alpha": 0.3}
    st.session_state["custom_settings"] = {"block_size": 16, "alpha": 1.0}
    assert W._resolve(None)[:2] == (16, 1.0)
    assert W._resolve({"block_size": 8})[:2] == (8, 0.1)       # explicit dict wins, missing keys -> constants
    st.session_state["custom_settings"] = {}                   # the embed page's initial value
    assert W.get_watermark_settings() == {"block_size": 8, "alpha": 0.1}


@pytest.mark.parametrize("name", ["pil_png_preserve1", "pil_png_preserve0"])
def test_resize_watermark_matches_reference(golden, name):
    g = golden(name)
    pr = name.endswith("1")
    png = g["png"].tobytes()
    got = W.resize_watermark(png, 16, 25, pr)
    assert got.mode == "L" and got.size == (25, 16)
    assert np.array_equal(np.array(got), g["wm"])
    got2 = W.resize_watermark(Image.open(io.BytesIO(png)).convert("RGB"), 16, 25, pr)
    assert np.array_equal(np.array(got2), g["wm"])
    assert np.array_equal(np.array(got), np.array(O.resize_watermark(png, 16, 25, pr)))


def test_resize_watermark_shapes():
    wide = Image.fromarray(np.zeros((10, 40), np.uint8))
    out = np.array(W.resize_watermark(wide, 20, 20, True))
    assert out.shape == (20, 20) and (out[:7] == 255).all() and (out[-7:] == 255).all() and (out[8:12] == 0).all()
    same = Image.fromarray(np.arange(12, dtype=np.uint8).reshape(3, 4))
    assert np.array_equal(np.array(W.resize_watermark(same, 3, 4, False)), np.array(same))


def test_shard_ranges_partition_by_image():
    for n in (0, 1, 7, 8, 1024, 1000):
        for parts in (1, 2, 4, 8):
            r = shard_ranges(n, parts)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - s for s, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_watermark_map_cache_and_decoding_helper(golden):
    g = golden("pil_png_preserve1")
    png = g["png"].tobytes()
    W.clear_watermark_cache()
    a = W.watermark_map(png, 16, 25, True)
    assert np.array_equal(a, g["wm"])
    assert W.watermark_map(png, 16, 25, True) is a                     # second call: cached object
    assert W.watermark_map(bytearray(png), 16, 25, True) is a
    b = W.watermark_map(png, 16, 25, False)                            # different key
    assert b is not a and np.array_equal(b, np.array(O.resize_watermark(png, 16, 25, False)))
    pil = Image.open(io.BytesIO(png))
    c = W.watermark_map(pil, 16, 25, True)                             # PIL input: computed, not cached
    assert c is not a and np.array_equal(c, a)
    big = W.prepare_for_decoding(Image.fromarray(a), scale=4, border=16)
    assert big.mode == "L" and big.size == (25 * 4 + 32, 16 * 4 + 32)
    assert set(np.unique(np.array(big))) <= {0, 255}


def test_hot_kernels_keep_their_matrices_in_registers():
    """ptxas resource usage of the built library (thatsmyface_b200.build keeps the -v logs): the tile kernel and the
    generic-N kernels up to 12 must have no stack frame.  A rolled loop once indexed the Gram matrix dynamically and
    pinned it to local memory (153 STL per block on the hot path, block sizes 10-16 at 0.3-0.4 of the roofline)."""
    from thatsmyface_b200 import build as B
    rows = B.resource_report()
    if not rows:
        pytest.skip("no ptxas logs beside the objects (library built elsewhere)")
    by_name = {r[0]: r for r in rows}
    tile = [r for n, r in by_name.items() if n.startswith("k_embed_tile<")]
    assert tile and all(r[2] == 0 and r[1] <= 128 for r in tile), tile
    for n, r in by_name.items():
        if n.startswith(("k_embed_fast_n<", "k_extract_fast_n<", "k_sigma0_fast_n<")):
            size = int(n.split("<")[1].split(",")[0])
            if size <= 12:
                assert r[2] == 0, f"{n}: {r[2]} bytes of stack frame"
        if n in ("k_embed_fast<8>", "k_sigma0_fast<8>"):
            assert r[2] == 0, (n, r)


def test_division_by_255_without_a_division_is_exact_for_every_byte():
    """tmf::modulate_sigma0 computes wm / 255.0 as q + fma(-q, 255, k) * r with r = RN(1/255), q = RN(k r).
    Emulated here with exact rationals (an fma is one rounding of the exact value): it must equal IEEE k / 255.0
    for all 256 bytes, and the test must be able to fail (the plain product is wrong for some bytes)."""
    from fractions import Fraction as F
    r = 1.0 / 255.0
    fma = lambda a, b, c: float(F(a) * F(b) + F(c))
    plain_wrong = 0
    for k in range(256):
        kd = float(k)
        q = kd * r
        got = fma(fma(-q, 255.0, kd), r, q)
        assert got == kd / 255.0, k
        plain_wrong += q != kd / 255.0
    assert plain_wrong > 0
