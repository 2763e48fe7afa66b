"""End-to-end (pinned host -> H2D -> kernel -> D2H -> pinned host) throughput of the batch
API for a few chunk sizes / pipeline depths.  python profiles/e2e_sweep.py [images]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from thatsmyface_b200 import pipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
imgs = torch.empty((n, bench.H, bench.W, 3), dtype=torch.uint8, device=dev)
bench.fill_images_device(imgs, 0, 17)
host_in = torch.empty((n, bench.H, bench.W, 3), dtype=torch.uint8, pin_memory=True)
host_in.copy_(imgs)
host_out = torch.empty_like(host_in).pin_memory()
wm = bench.make_wm_map()
for depth in (2, 3, 4):
    for mb in (16, 32, 64, 96, 192):
        pipeline.run_batch("embed", host_in, None, wm, 0.1, 8, 1, [0], host_out, chunk_bytes=mb << 20, depth=depth)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            pipeline.run_batch("embed", host_in, None, wm, 0.1, 8, 1, [0], host_out, chunk_bytes=mb << 20, depth=depth)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"depth {depth} chunk {mb:4d} MB: {n * bench.PX / dt / 1e6:9.0f} MP/s  ({2 * n * bench.PX * 3 / dt / 1e9:.1f} GB/s PCIe both ways)", flush=True)
