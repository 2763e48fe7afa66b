"""Small all-kernels exercise for compute-sanitizer (memcheck): odd sizes, strips, every
alignment path, per-image maps, every mode, the TMA-tiled kernel, the other block sizes, the map kernel,
the host-buffer context, taps.  Run plain first, then under the tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from thatsmyface_b200 import watermarking as W

rng = np.random.default_rng(0)
for (h, w) in ((64, 64), (70, 93), (40, 52), (40, 51), (8, 8), (5, 7), (24, 1000), (24, 128), (40, 384)):
    n = 3
    x = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
    for shared in (True, False):
        shape = (h // 8, w // 8) if shared else (n, h // 8, w // 8)
        m = torch.from_numpy(rng.integers(0, 256, shape, dtype=np.uint8)).cuda()
        for mode in (0, 1, 2):
            o = W.embed_tensor(x, m, 0.1, 8, mode)
            e = W.extract_tensor(o, x, 0.1, 8, mode)
            s = W.sigma0_tensor(x, 8, mode)
    # an unaligned view (base pointer + 1 byte)
    flat = torch.empty(n * h * w * 3 + 8, dtype=torch.uint8, device="cuda")
    v = flat[1:1 + n * h * w * 3].view(n, h, w, 3)
    v.copy_(x)
    assert v.data_ptr() % 2 == 1
    lib_out = torch.empty_like(x)
    from thatsmyface_b200 import _lib
    m = torch.from_numpy(rng.integers(0, 256, (h // 8, w // 8), dtype=np.uint8)).cuda()
    for mode in (0, 1):
        _lib.check(_lib.load().tmf_embed_rgb8(v.data_ptr(), lib_out.data_ptr(), n, h, w, h * w * 3, m.data_ptr(), 1,
                                              0.1, 8, mode, torch.cuda.current_stream().cuda_stream))
        assert torch.equal(lib_out, W.embed_tensor(x, m, 0.1, 8, mode))
# the other block sizes: aligned and ragged shapes, both modes (generic-N fast kernels, shared-memory faithful kernels)
for bs in (4, 6, 10, 12, 14, 16):
    for (h, w) in ((3 * bs, 5 * bs + 2), (4 * bs, 32 * bs), (2 * bs + 1, 7 * bs + 3)):
        x = torch.from_numpy(rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)).cuda()
        m = torch.from_numpy(rng.integers(0, 256, (2, h // bs, w // bs), dtype=np.uint8)).cuda()
        for mode in (0, 1):
            o = W.embed_tensor(x, m, 0.1, bs, mode)
            W.extract_tensor(o, x, 0.1, bs, mode)
            W.sigma0_tensor(x, bs, mode)
# watermark map on the device and the host-buffer pipeline
src = torch.from_numpy(rng.integers(0, 256, (3, 290, 290), dtype=np.uint8)).cuda()
W.watermark_map_tensor(src, 135, 240, True)
W.watermark_map_tensor(src, 36, 64, False)
imgs = rng.integers(0, 256, (5, 72, 128, 3), dtype=np.uint8)
wmh = rng.integers(0, 256, (9, 16), dtype=np.uint8)
outs = W.embed_watermark_batch(imgs, wmh, 0.1, 8, 1)
W.extract_watermark_batch(outs, imgs, 0.1, 8, 1)
D = torch.from_numpy(rng.normal(size=(1000, 8, 8)).astype(np.float32)).cuda()
D[5] = 0
W.svd8x8(D, vectors=True, complete_u=True)
W.svd8x8(D[:130], vectors=False)
W.dct8x8(D[:129])
W.dct8x8(D[:129], inverse=True)
W.rgb_to_ycbcr(rng.integers(0, 256, (33, 17, 3), dtype=np.uint8))
W.ycbcr_to_rgb(rng.random((33, 17, 3), dtype=np.float32))
torch.cuda.synchronize()
print("sanitize_smoke ok")
