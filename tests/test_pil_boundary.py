"""The single-image API's PIL boundary (thatsmyface_b200/watermarking.py): PIL's 4-byte "RGB" storage
crosses PCIe as it is (Arrow export / Image.fromarrow) and is packed / unpacked on the device
(tmf_rgbx8_to_rgb8 / tmf_rgb8_to_rgbx8).  Host helpers are checked here without a GPU, the kernels and
the equivalence with the packed-bytes route on the B200."""
import numpy as np
import pytest
from PIL import Image

from thatsmyface_b200 import _lib
from thatsmyface_b200 import build as tmf_build
from thatsmyface_b200 import watermarking as W

tmf_build.build()          # no-op when the in-tree .so is current

needs_arrow = pytest.mark.skipif(not W._fast_pil(), reason="pyarrow / Pillow Arrow interface not available")
SHAPES = [(1, 1), (7, 5), (64, 64), (135, 241), (1080, 1920)]


@needs_arrow
@pytest.mark.parametrize("shape", SHAPES)
def test_stage_view_is_pils_storage_and_is_reusable(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    st = W._stage_acquire((w, h), pin=False)
    for _ in range(2):
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        st.fill(Image.fromarray(a))
        assert st.view.shape == (h, w, 4) and (st.view[..., :3] == a).all()
    W._stage_release(st)
    again = W._stage_acquire((w, h), pin=False)
    assert again is st                                    # the pool hands the same block out again
    other = W._stage_acquire((w, h), pin=False)
    assert other is not st                                # ... but never twice at the same time
    W._stage_release(again)
    W._stage_release(other)


@needs_arrow
def test_stage_pool_is_bounded(monkeypatch):
    monkeypatch.setattr(W, "_STAGE_POOL_BYTES", 64 * 64 * 4 * 3)
    held = [W._stage_acquire((64, 64), pin=False) for _ in range(8)]
    for st in held:
        W._stage_release(st)
    assert W._stage_free_bytes <= W._STAGE_POOL_BYTES
    assert sum(len(v) for v in W._stage_free.values()) <= 3 + 5      # other sizes may be pooled by other tests


@needs_arrow
@pytest.mark.parametrize("shape", SHAPES)
def test_output_wrap_is_a_real_rgb_image(shape):
    h, w = shape
    a = np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)
    x4 = np.concatenate([a, np.full((h, w, 1), 255, np.uint8)], axis=2)
    im = W._rgbx_array_to_pil(x4)
    del x4
    assert im.mode == "RGB" and im.size == (w, h) and (np.asarray(im) == a).all()
    im2 = im.copy()
    im2.putpixel((0, 0), (1, 2, 3))                      # copy-on-write: the wrapped buffer is not the caller's problem
    assert im2.getpixel((0, 0)) == (1, 2, 3) and im.convert("L").size == (w, h)


def test_format_taps_validate_arguments_without_a_gpu():
    lib = _lib.load()
    assert lib.tmf_rgbx8_to_rgb8(None, None, 0, None) == 0
    assert lib.tmf_rgbx8_to_rgb8(None, None, -1, None) == _lib.ERR_BAD_ARG
    assert lib.tmf_rgb8_to_rgbx8(None, None, 5, 255, None) == _lib.ERR_BAD_ARG
    assert lib.tmf_rgbx8_to_rgb8(16, 2, 4, None) == _lib.ERR_BAD_ARG          # rgb not 4-byte aligned


@pytest.mark.gpu
@pytest.mark.parametrize("npx", [1, 3, 4, 5, 1023, 4096, 1920 * 1080 + 2])
def test_format_kernels_against_numpy(npx):
    import torch

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(npx)
    x4 = torch.randint(0, 256, (npx, 4), dtype=torch.uint8, device="cuda", generator=g)
    x3 = torch.empty((npx, 3), dtype=torch.uint8, device="cuda")
    _lib.check(lib.tmf_rgbx8_to_rgb8(x4.data_ptr(), x3.data_ptr(), npx, None))
    assert torch.equal(x3, x4[:, :3])
    y4 = torch.zeros((npx, 4), dtype=torch.uint8, device="cuda")
    _lib.check(lib.tmf_rgb8_to_rgbx8(x3.data_ptr(), y4.data_ptr(), npx, 255, None))
    assert torch.equal(y4[:, :3], x3) and bool((y4[:, 3] == 255).all())


@pytest.mark.gpu
@needs_arrow
@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "P"])
def test_both_boundary_routes_give_the_same_tensors_and_images(mode):
    import torch

    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (123, 250, 3), dtype=np.uint8)
    im = Image.fromarray(a).convert(mode)
    ref = W._to_device(W._pil_to_rgb_array(im))
    for _ in range(2):                                    # second pass reuses the staging block
        got = W._pil_to_device_rgb(im)
        assert torch.equal(got, ref)
    out = W._device_rgb_to_pil(ref)
    assert out.mode == "RGB" and (np.asarray(out) == ref.cpu().numpy()).all()
