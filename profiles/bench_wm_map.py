"""Measurement of the watermark-map row: tmf_wm_map_l8 on one B200 with
the reference's own call for the same work (PIL LANCZOS + paste, modules/watermarking.py:105-123)
timed beside it on the host.  Algorithmic bytes: src_h*src_w read + target_h*target_w written per
map.  Prints one JSON line per configuration.

    python profiles/bench_wm_map.py > gpurun_out/wm_map.jsonl
"""
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thatsmyface_b200 import watermarking as W  # noqa: E402


def qr_like(n, size, seed):
    rng = np.random.default_rng(seed)
    cells = (rng.integers(0, 2, (n, 45, 45)) * 255).astype(np.uint8)
    rep = -(-size // 45)
    return np.ascontiguousarray(np.kron(cells, np.ones((1, rep, rep), np.uint8))[:, :size, :size])


def main():
    peak = 6542.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for n, (th, tw) in ((1, (135, 240)), (64, (135, 240)), (1024, (135, 240)), (256, (270, 480))):
        src = qr_like(n, 1000, n)
        t = torch.from_numpy(src).cuda()
        out = torch.empty((n, th, tw), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            W.watermark_map_tensor(t, th, tw, True, out=out)
        torch.cuda.synchronize()
        reps = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            W.watermark_map_tensor(t, th, tw, True, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        # host: the reference's call on already decoded mode-L images, one core
        k = min(n, 16)
        imgs = [Image.fromarray(src[i], "L") for i in range(k)]
        t0 = time.perf_counter()
        ref = [np.array(W.resize_watermark(im, th, tw, True)) for im in imgs]
        cpu_ms = (time.perf_counter() - t0) * 1e3 / k
        same = bool(np.array_equal(out[:k].cpu().numpy(), np.stack(ref)))
        gb = n * (1000 * 1000 + th * tw) / 1e9
        print(json.dumps({"op": "wm_map_l8", "n": n, "src": [1000, 1000], "target": [th, tw], "preserve_ratio": True,
                          "ms_per_call": round(ms, 4), "maps_per_s": round(n / ms * 1e3, 1),
                          "algorithmic_GBps": round(gb / ms * 1e3, 1), "hbm_frac": round(gb / ms * 1e3 / peak, 4),
                          "pil_ms_per_map_1core": round(cpu_ms, 3), "identical_to_pil": same}), flush=True)


if __name__ == "__main__":
    main()
