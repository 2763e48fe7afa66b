"""Measurement of the page loop (caller side): N encoded uploads ->
embed_watermark -> PNG bytes, (a) the reference page's shape of loop with the drop-in function
(one image after the other), (b) page_loop.embed_watermark_many on the lane pool.  One JSON line
per configuration.

    python profiles/bench_page_loop.py > gpurun_out/page_loop.jsonl
"""
import io
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thatsmyface_b200 import page_loop  # noqa: E402
from thatsmyface_b200 import watermarking as W  # noqa: E402


def upload(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = 120 + 70 * np.sin(x / 97.0) * np.cos(y / 71.0)
    a = np.clip(base[..., None] + np.array([10.0, 0, -10.0]) + rng.normal(0, 8, (h, w, 3)), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="JPEG", quality=92)     # uploads are mostly JPEG photographs
    return buf.getvalue()


def qr_png():
    rng = np.random.default_rng(1)
    cells = (rng.integers(0, 2, (45, 45)) * 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(np.kron(cells, np.ones((23, 23), np.uint8))[:1000, :1000]).save(buf, format="PNG")
    return buf.getvalue()


def main():
    wm = qr_png()
    n = 30                                                      # MAX_IMAGES, constants.py:2
    files = [upload(1080, 1920, k) for k in range(n)]
    mp = n * 1080 * 1920 / 1e6
    W.embed_watermark(Image.open(io.BytesIO(files[0])), wm, True)   # warm-up: context, map cache
    torch.cuda.synchronize()

    for png in (False, True):
        t0 = time.perf_counter()
        seq = []
        for f in files:                                         # embed_watermark_page.py:492-545
            out = W.embed_watermark(Image.open(io.BytesIO(f)), wm, preserve_ratio=True)
            if png:
                b = io.BytesIO()
                out.save(b, format="PNG")
                seq.append((out, b.getvalue()))
            else:
                seq.append(out)
        t_seq = time.perf_counter() - t0
        for lanes in (2, 4, 8, 16):
            page_loop.embed_watermark_many(files[:lanes], wm, True, png=png, lanes=lanes)   # warm the lanes' pools
            t0 = time.perf_counter()
            got = page_loop.embed_watermark_many(files, wm, True, png=png, lanes=lanes)
            t_many = time.perf_counter() - t0
            a = got[-1][0] if png else got[-1]
            b = seq[-1][0] if png else seq[-1]
            print(json.dumps({"op": "embed page loop", "images": n, "size": [1080, 1920], "input": "JPEG bytes",
                              "output": "PIL + PNG bytes" if png else "PIL", "host_cores": os.cpu_count(),
                              "lanes": lanes, "sequential_s": round(t_seq, 3), "lanes_s": round(t_many, 3),
                              "sequential_MPps": round(mp / t_seq, 1), "lanes_MPps": round(mp / t_many, 1),
                              "speedup": round(t_seq / t_many, 2),
                              "identical": bool(np.array_equal(np.asarray(a), np.asarray(b)))}), flush=True)


if __name__ == "__main__":
    main()
