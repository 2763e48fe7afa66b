"""The pages' per-image loops on the lane pool (thatsmyface_b200/page_loop.py): same results as
the single-image drop-in functions, any sizes and modes, order preserved."""
import io

import numpy as np
import pytest
from PIL import Image

torch = pytest.importorskip("torch")


def _img(h, w, seed, mode="RGB"):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = 120 + 70 * np.sin(x / 37.0) * np.cos(y / 29.0)
    a = np.clip(base[..., None] + np.array([10.0, 0, -10.0]) + rng.normal(0, 8, (h, w, 3)), 0, 255).astype(np.uint8)
    im = Image.fromarray(a)
    return im if mode == "RGB" else im.convert(mode)


def _wm_png(seed=0, size=200):
    rng = np.random.default_rng(seed)
    cells = (rng.integers(0, 2, (25, 25)) * 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(np.kron(cells, np.ones((size // 25, size // 25), np.uint8))).save(buf, format="PNG")
    return buf.getvalue()


def test_without_a_gpu_the_loop_raises_like_the_single_image_call():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from thatsmyface_b200 import page_loop

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        page_loop.embed_watermark_many([_img(16, 16, 0)], _wm_png())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        page_loop.extract_watermark_many([(_img(16, 16, 0), _img(16, 16, 0))])


@pytest.mark.gpu
def test_embed_many_equals_the_single_image_function_on_mixed_sizes_and_modes():
    from thatsmyface_b200 import page_loop
    from thatsmyface_b200 import watermarking as W

    wm = _wm_png()
    shapes = [(240, 320, "RGB"), (123, 211, "RGB"), (480, 640, "RGBA"), (64, 64, "L"), (301, 97, "P"),
              (240, 320, "RGB"), (1080, 1920, "RGB"), (8, 8, "RGB"), (77, 500, "RGB")] * 2
    images = [_img(h, w, k, m) for k, (h, w, m) in enumerate(shapes)]
    ref = [W.embed_watermark(im, wm, preserve_ratio=True) for im in images]
    got = page_loop.embed_watermark_many(images, wm, preserve_ratio=True, lanes=4)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g.mode == "RGB" and g.size == r.size
        assert np.array_equal(np.asarray(g), np.asarray(r))


@pytest.mark.gpu
def test_embed_many_png_bytes_files_and_settings():
    from thatsmyface_b200 import page_loop
    from thatsmyface_b200 import watermarking as W

    wm = _wm_png(3)
    settings = {"block_size": 4, "alpha": 0.5}
    encoded = []
    for k in range(6):
        buf = io.BytesIO()
        _img(100 + 8 * k, 140, k).save(buf, format="PNG")
        encoded.append(buf.getvalue())
    items = [encoded[0], io.BytesIO(encoded[1])] + encoded[2:]        # bytes and file objects
    got = page_loop.embed_watermark_many(items, wm, True, settings, png=True, lanes=3,
                                         png_options={"compress_level": 1})
    for (im, png), enc in zip(got, encoded):
        ref = W.embed_watermark(Image.open(io.BytesIO(enc)), wm, True, settings)
        assert np.array_equal(np.asarray(im), np.asarray(ref))
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))), np.asarray(ref))


@pytest.mark.gpu
def test_extract_many_round_trip_and_errors():
    from thatsmyface_b200 import page_loop
    from thatsmyface_b200 import watermarking as W

    wm = _wm_png(5)
    originals = [_img(256 + 16 * k, 384, 10 + k) for k in range(5)]
    marked = page_loop.embed_watermark_many(originals, wm, True)
    maps = page_loop.extract_watermark_many(list(zip(marked, originals)), lanes=2)
    for m, w_, o in zip(maps, marked, originals):
        ref = W.extract_watermark(w_, o)
        assert m.mode == "L" and m.size == ref.size and np.array_equal(np.asarray(m), np.asarray(ref))
        target = np.asarray(W.resize_watermark(wm, m.size[1], m.size[0], True))
        decided = (target < 64) | (target > 192)
        assert ((np.asarray(m) >= 128) == (target >= 128))[decided].mean() > 0.99
    with pytest.raises(ValueError, match="same size"):
        page_loop.extract_watermark_many([(originals[0], originals[1])])
    with pytest.raises(ValueError, match="not supported"):
        page_loop.embed_watermark_many(originals, wm, True, {"block_size": 7, "alpha": 0.1})


@pytest.mark.gpu
def test_generator_preserves_order_and_streams_results():
    from thatsmyface_b200 import page_loop

    wm = _wm_png(7)
    sizes = [(64 + 8 * (k % 5), 96 + 8 * (k % 3)) for k in range(23)]
    it = page_loop.embed_watermark_iter((_img(h, w, k) for k, (h, w) in enumerate(sizes)), wm, lanes=4)
    seen = [im.size for im in it]
    assert seen == [(w, h) for h, w in sizes]


@pytest.mark.gpu
def test_lanes_over_two_devices():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from thatsmyface_b200 import page_loop
    from thatsmyface_b200 import watermarking as W

    wm = _wm_png(9)
    images = [_img(200, 300, k) for k in range(8)]
    got = page_loop.embed_watermark_many(images, wm, True, lanes=4, devices=[0, 1])
    for g, im in zip(got, images):
        assert np.array_equal(np.asarray(g), np.asarray(W.embed_watermark(im, wm, True)))
