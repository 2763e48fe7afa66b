"""The TMA-tiled persistent embed kernel (k_embed_tile, the default for 16-byte aligned batches whose
rows hold a multiple of 16 blocks) against the per-thread kernel on the same pixels: the outputs
must be identical (tests/tools/ab_tma.py: aligned buffers take the tile kernel, the same data at an
8-byte offset cannot).  Extract has one kernel; the tool checks it is insensitive to the offset too."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1080, 1920),     # 15 boxes per block-row: every other tile wraps to the next row
                                   (3, 264, 512),       # 4 boxes per row
                                   (1, 8, 128),         # one box: half a tile
                                   (5, 24, 384),        # 3 boxes per row, 9 per image: tiles straddle images
                                   (37, 2160, 3840)])   # more tiles than resident warps: the persistent loop
def test_tma_kernel_equals_per_thread_kernel(shape):
    n, h, w = shape
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ab_tma.py"), str(n), str(h), str(w)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    counts = [int(m) for m in re.findall(r"mismatching samples (\d+) of", r.stdout)]
    assert len(counts) == 3 and counts == [0, 0, 0], r.stdout
    m = re.search(r"extract paths \([01], 0\): mismatching levels (\d+) of", r.stdout)
    assert m and int(m.group(1)) == 0, r.stdout


@pytest.mark.gpu
def test_tile_kernel_per_image_maps_strips_and_padded_image_stride():
    """In-process: per-image maps, an H % 8 strip under the tiles, and (through the C ABI) an image
    stride larger than one image - each against the per-thread kernel on an 8-byte-offset copy."""
    import numpy as np
    import torch

    from thatsmyface_b200 import _lib
    from thatsmyface_b200 import watermarking as W

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(3)
    n, h, w = 3, 45, 256                                     # 5 block-rows + a 5-row strip; 2 boxes per block-row
    imgs = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    wms = torch.randint(0, 256, (n, h // 8, w // 8), dtype=torch.uint8, device="cuda", generator=g)
    wms[:, :, ::3] = 0

    def offset_copy(t):
        raw = torch.empty(t.numel() + 16, dtype=torch.uint8, device="cuda")
        v = raw[8:8 + t.numel()].view(t.shape)
        v.copy_(t)
        return v

    a = W.embed_tensor(imgs, wms, 0.1, 8, 1)
    assert lib.tmf_last_fast_path() == 1
    b = W.embed_tensor(offset_copy(imgs), wms, 0.1, 8, 1, out=offset_copy(torch.zeros_like(imgs)))
    assert lib.tmf_last_fast_path() == 0
    assert torch.equal(a, b)

    # padded image stride (a multiple of 16: still the tile kernel), shared map, straight through the C ABI
    img_bytes = h * w * 3
    stride = img_bytes + 48
    src = torch.zeros((n, stride), dtype=torch.uint8, device="cuda")
    dst = torch.full((n, stride), 7, dtype=torch.uint8, device="cuda")
    src[:, :img_bytes] = imgs.view(n, -1)
    wm0 = wms[0].contiguous()
    _lib.check(lib.tmf_embed_rgb8(src.data_ptr(), dst.data_ptr(), n, h, w, stride, wm0.data_ptr(), 1, 0.1, 8, 1,
                                  torch.cuda.current_stream().cuda_stream))
    assert lib.tmf_last_fast_path() == 1
    want = W.embed_tensor(offset_copy(imgs), wm0, 0.1, 8, 1, out=offset_copy(torch.zeros_like(imgs)))
    assert torch.equal(dst[:, :img_bytes].view(n, h, w, 3), want)
    assert int(dst[:, img_bytes:].min()) == 7 and int(dst[:, img_bytes:].max()) == 7      # the padding is not touched
    # a stride that is not a multiple of 16 takes the per-thread kernel, same pixels
    stride2 = img_bytes + 8
    src2 = torch.zeros((n, stride2), dtype=torch.uint8, device="cuda")
    dst2 = torch.zeros((n, stride2), dtype=torch.uint8, device="cuda")
    src2[:, :img_bytes] = imgs.view(n, -1)
    _lib.check(lib.tmf_embed_rgb8(src2.data_ptr(), dst2.data_ptr(), n, h, w, stride2, wm0.data_ptr(), 1, 0.1, 8, 1,
                                  torch.cuda.current_stream().cuda_stream))
    assert lib.tmf_last_fast_path() == 0
    assert torch.equal(dst2[:, :img_bytes].view(n, h, w, 3), want)
