"""Out-of-bounds WRITES, checked without compute-sanitizer (closed on the GPU pool): every output lives inside a
larger canary-filled allocation, and after the kernel every byte outside the image / map must still be the canary.
Covers the paths that write wider than a block row (16-byte rows of block size 16, aligned words shared between
neighbouring blocks for sizes 6 / 10 / 14, TMA box stores, strips) at aligned and unaligned offsets, in every mode.
Reads past the end cannot be seen this way; the inputs are placed at the very END of their allocations' used part
with canaries behind them, and results are compared with the same call on a roomy copy (a read of canary bytes that
influenced a pixel would change it)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from thatsmyface_b200 import _lib  # noqa: E402
from thatsmyface_b200 import watermarking as W  # noqa: E402

CANARY = 0xA5
PAD = 256


def _embedded(nbytes, offset):
    """A canary-filled CUDA buffer and the view of `nbytes` bytes at PAD + offset inside it."""
    buf = torch.full((PAD + offset + nbytes + PAD,), CANARY, dtype=torch.uint8, device="cuda")
    return buf, buf[PAD + offset: PAD + offset + nbytes]


def _outside_intact(buf, offset, nbytes):
    head = buf[: PAD + offset]
    tail = buf[PAD + offset + nbytes:]
    return bool((head == CANARY).all()) and bool((tail == CANARY).all())


SHAPES = {
    8: [(16, 128), (24, 384), (17, 131), (8, 8), (40, 52)],          # tile path (w % 128 == 0), strips, ragged
    4: [(8, 64), (9, 67)], 6: [(12, 60), (13, 62), (12, 54)], 10: [(20, 80), (21, 93), (20, 90)],
    12: [(24, 96), (25, 99)], 14: [(28, 112), (29, 115), (28, 126)], 16: [(32, 128), (33, 131), (32, 144)],
}


@pytest.mark.parametrize("offset", [0, 1, 2, 4, 16])
@pytest.mark.parametrize("bs", [4, 6, 8, 10, 12, 14, 16])
def test_embed_and_extract_write_nothing_outside_their_outputs(bs, offset):
    rng = np.random.default_rng(bs * 31 + offset)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    for (h, w) in SHAPES[bs]:
        n = 2
        nbytes = n * h * w * 3
        nbh, nbw = h // bs, w // bs
        x_np = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        inbuf, xin = _embedded(nbytes, offset)
        xin.copy_(torch.from_numpy(x_np).reshape(-1))
        wm = torch.from_numpy(rng.integers(0, 256, (n, max(nbh, 1), max(nbw, 1)), dtype=np.uint8)).cuda()
        roomy = torch.from_numpy(x_np).cuda()
        for mode in ((0, 1, 2) if bs == 8 else (0, 1)):
            obuf, out = _embedded(nbytes, offset)
            _lib.check(lib.tmf_embed_rgb8(xin.data_ptr(), out.data_ptr(), n, h, w, h * w * 3, wm.data_ptr(), 0, 0.1, bs, mode, st))
            torch.cuda.synchronize()
            assert _outside_intact(obuf, offset, nbytes), f"embed bs {bs} mode {mode} {h}x{w} +{offset}: wrote outside the output"
            assert _outside_intact(inbuf, offset, nbytes), "the input's surroundings changed"
            ref = W.embed_tensor(roomy, wm[:, :nbh, :nbw].contiguous() if nbh and nbw else wm[:, :0, :0].contiguous(), 0.1, bs, mode)
            assert torch.equal(out.view(n, h, w, 3), ref), f"embed bs {bs} mode {mode} {h}x{w} +{offset}: result depends on the placement"
            if nbh and nbw:
                mbytes = n * nbh * nbw
                ebuf, ext = _embedded(mbytes, offset)
                _lib.check(lib.tmf_extract_rgb8(out.data_ptr(), xin.data_ptr(), ext.data_ptr(), n, h, w, h * w * 3, 0.1, bs, mode, st))
                torch.cuda.synchronize()
                assert _outside_intact(ebuf, offset, mbytes), f"extract bs {bs} mode {mode} {h}x{w} +{offset}: wrote outside the map"
                assert torch.equal(ext.view(n, nbh, nbw), W.extract_tensor(ref, roomy, 0.1, bs, mode))


def test_wm_map_and_taps_write_nothing_outside_their_outputs():
    rng = np.random.default_rng(5)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    src = torch.from_numpy(rng.integers(0, 256, (3, 97, 113), dtype=np.uint8)).cuda()
    for (th, tw, pr) in ((13, 24, 1), (13, 24, 0), (7, 5, 1)):
        mbytes = 3 * th * tw
        mbuf, maps = _embedded(mbytes, 3)
        ws_bytes = lib.tmf_wm_map_workspace_bytes(3, 97, 113, th, tw, pr)
        wbuf = torch.full((ws_bytes + 2 * PAD,), CANARY, dtype=torch.uint8, device="cuda")
        base = wbuf.data_ptr() + PAD
        base += (-base) % 16
        _lib.check(lib.tmf_wm_map_l8(src.data_ptr(), 3, 97, 113, 97 * 113, maps.data_ptr(), th, tw, pr, base, ws_bytes, st))
        torch.cuda.synchronize()
        assert _outside_intact(mbuf, 3, mbytes)
        off = base - wbuf.data_ptr()
        assert bool((wbuf[:off] == CANARY).all()) and bool((wbuf[off + ws_bytes:] == CANARY).all()), "workspace overrun"
        assert torch.equal(maps.view(3, th, tw), W.watermark_map_tensor(src, th, tw, bool(pr)))
