"""A/B check of the two FAST block-8 embed code paths on the same pixels: 16-byte aligned buffers take
the TMA-tiled persistent kernel (k_embed_tile), the same data at an 8-byte offset takes the per-thread
kernel.  Outputs must be identical.
Usage: python tests/tools/ab_tma.py [n] [h] [w]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
from thatsmyface_b200 import _lib
from thatsmyface_b200 import watermarking as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h = int(sys.argv[2]) if len(sys.argv) > 2 else bench.H
w = int(sys.argv[3]) if len(sys.argv) > 3 else bench.W
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
imgs = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev, generator=g)
imgs[n // 2:, : h // 2] //= 8           # dark region: zero-ish blocks, clipping
wm = torch.randint(0, 256, (h // 8, w // 8), dtype=torch.uint8, device=dev, generator=g)
wm[:, : w // 32] = 0                     # zero-mark lanes
nbytes = imgs.numel()
raw_in = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
raw_out = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
off_in = raw_in[8:8 + nbytes].view(n, h, w, 3)
off_out = raw_out[8:8 + nbytes].view(n, h, w, 3)
off_in.copy_(imgs)
assert imgs.data_ptr() % 16 == 0 and off_in.data_ptr() % 16 == 8
lib = _lib.load()
for rep in range(3):
    a = W.embed_tensor(imgs, wm, 0.1, 8, 1)
    path_a = lib.tmf_last_fast_path()
    W.embed_tensor(off_in, wm, 0.1, 8, 1, out=off_out)
    path_b = lib.tmf_last_fast_path()
    torch.cuda.synchronize()
    assert (path_a, path_b) == (1, 0), f"expected the tile kernel then the per-thread kernel, got {(path_a, path_b)}"
    d = (a != off_out)
    bad = int(d.sum())
    print(f"rep {rep}: mismatching samples {bad} of {a.numel()}")
    if bad:
        idx = d.nonzero()[:12].tolist()
        for (i, y, x, c) in idx:
            gb = (y // 8) * (w // 8) + x // 8
            print(f"  img {i} y {y} x {x} c {c}: tma {int(a[i, y, x, c])} ref {int(off_out[i, y, x, c])}  block-in-img {gb} lane {gb % 32} by {y//8} bx {x//8}")
        per_img = d.flatten(1).sum(1).tolist()
        print("  per image:", per_img)
        rows = d.any(dim=3).any(dim=2).sum(1).tolist()
        print("  rows touched per image:", rows)

# extract (one kernel): aligned buffers against the offset copies of the embedded images
ea = W.extract_tensor(a, imgs, 0.1, 8, 1)
path_a = lib.tmf_last_fast_path()
eb = W.extract_tensor(off_out, off_in, 0.1, 8, 1)
path_b = lib.tmf_last_fast_path()
torch.cuda.synchronize()
print(f"extract paths {(path_a, path_b)}: mismatching levels {int((ea != eb).sum())} of {ea.numel()}")
