import sys, time, cProfile, pstats, io
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
from PIL import Image
import bench
from thatsmyface_b200 import watermarking as Wm
h, w = 2160, 3840
x = torch.empty((1, h, w, 3), dtype=torch.uint8, device="cuda"); bench.fill_images_device(x, 0, 3)
img = Image.fromarray(x[0].cpu().numpy())
payload, png = bench.payload_and_png()
s = {"block_size": 8, "alpha": 0.1, "mode": 1}
for _ in range(3): out = Wm.embed_watermark(img, png, True, s)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); out = Wm.embed_watermark(img, png, True, s); ts.append(time.perf_counter() - t0)
print("embed ms", [round(t*1e3,2) for t in ts])
pr = cProfile.Profile(); pr.enable()
for _ in range(5): out = Wm.embed_watermark(img, png, True, s)
pr.disable(); st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(18); print(st.getvalue()[:3500])
# with big pinned buffers alive, as in bench.py
big = torch.empty((256, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); out = Wm.embed_watermark(img, png, True, s); ts.append(time.perf_counter() - t0)
print("embed ms with 1.6 GB pinned alive", [round(t*1e3,2) for t in ts])
