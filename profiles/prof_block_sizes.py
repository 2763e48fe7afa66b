"""One embed launch per block size on 64 x 1080p, for ncu (`-k regex:k_embed_fast_n`).  Usage: prof_block_sizes.py 10 12 16"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from thatsmyface_b200 import watermarking as W

n = 64
imgs = torch.empty((n, bench.H, bench.W, 3), dtype=torch.uint8, device="cuda")
bench.fill_images_device(imgs, 0, 17)
out = torch.empty_like(imgs)
for bs in [int(a) for a in sys.argv[1:]] or [10, 12, 16]:
    wm = (torch.rand((bench.H // bs, bench.W // bs), device="cuda") < 0.5).to(torch.uint8) * 255
    W.embed_tensor(imgs, wm, 0.1, bs, 1, out=out)
    torch.cuda.synchronize()
    print("bs", bs, "done")
