"""Small driver for ncu: a few launches of each fused kernel on a 1080p sub-batch.
Usage: python profiles/prof_kernels.py [images]   (run plain first, then under ncu)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from thatsmyface_b200 import watermarking as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
imgs = torch.empty((n, bench.H, bench.W, 3), dtype=torch.uint8, device=dev)
bench.fill_images_device(imgs, 0, 17)
wm = torch.from_numpy(bench.make_wm_map()).to(dev)
out = torch.empty_like(imgs)
for rep in range(2):
    for mode in (1, 0):
        W.embed_tensor(imgs, wm, 0.1, 8, mode, out=out)
        W.extract_tensor(out, imgs, 0.1, 8, mode)
torch.cuda.synchronize()
print("ok")
