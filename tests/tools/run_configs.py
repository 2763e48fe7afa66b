"""BASELINE.json configs 1, 2, 4 and 5 on one B200, for the record (bench.py is config 3).
Lives under tests/ because it uses the oracle as its checker and CPU timing reference.
Writes one JSON document; run on the GPU box:  python tests/tools/run_configs.py > gpurun_out/rNN_configs.json"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # qr_util

import numpy as np
import torch
from PIL import Image

import bench
import qr_util as Q
from oracle import wm_oracle as O            # checker / CPU timing only
from thatsmyface_b200 import watermarking as W


def natural(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = (120 + 70 * np.sin(x / 97.0) * np.cos(y / 71.0))[..., None] + np.array([10.0, 0.0, -10.0]) + rng.normal(0, 8, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def median_ms(fn, reps, sync=True):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        if sync:
            torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def kernel_us(fn, reps=100):
    for _ in range(5):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs) * 1e3


def main():
    res = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    # ---- config 1: 512x512, text-derived QR, reference CPU path next to it
    rgb = natural(512, 512, 2)
    img = Image.fromarray(rgb)
    png = Q.qr_png(Q.encrypt("hello"))
    c1 = {}
    for mode, name in ((1, "fast"), (0, "faithful")):
        s = {"block_size": 8, "alpha": 0.1, "mode": mode}
        out = W.embed_watermark(img, png, True, s)
        ext = W.extract_watermark(out, img, s)
        c1[name] = {"embed_ms_e2e_pil": round(median_ms(lambda: W.embed_watermark(img, png, True, s), 20), 3),
                    "extract_ms_e2e_pil": round(median_ms(lambda: W.extract_watermark(out, img, s), 20), 3),
                    "payload_decoded": Q.decode_map(np.array(ext)) is not None}
    wm = np.array(O.resize_watermark(png, 64, 64, True))
    t0 = time.perf_counter()
    ref = O.embed_array(rgb, wm, style="loop")
    t1 = time.perf_counter()
    ref_ext = O.extract_array(ref, rgb, style="loop")
    t2 = time.perf_counter()
    gpu_out = np.array(W.embed_watermark(img, png, True, {"block_size": 8, "alpha": 0.1, "mode": 0}))
    c1["cpu_reference_port_loop_form"] = {"embed_s": round(t1 - t0, 3), "extract_s": round(t2 - t1, 3),
                                          "embed_MPps": round(0.262144 / (t1 - t0), 4), "cores": 1}
    c1["parity_vs_oracle"] = {"max_abs_pixel_diff": int(np.abs(gpu_out.astype(int) - ref.astype(int)).max()),
                              "fraction_differing": float((gpu_out != ref).mean()),
                              "payload_equal": Q.decode_map(ref_ext) == Q.decode_map(np.array(
                                  W.extract_watermark(Image.fromarray(gpu_out), img, {"block_size": 8, "alpha": 0.1, "mode": 0})))}
    res["config1_512x512"] = c1

    # ---- config 2: one 4K image, latency path
    rgb4 = natural(2160, 3840, 3)
    img4 = Image.fromarray(rgb4)
    png4 = Q.qr_png(Q.encrypt("Test" * 10))
    x = torch.from_numpy(rgb4).cuda()
    m = torch.from_numpy(np.array(W.resize_watermark(png4, 270, 480, True))).cuda()
    c2 = {}
    for mode, name in ((1, "fast"), (0, "faithful")):
        s = {"block_size": 8, "alpha": 0.1, "mode": mode}
        o = W.embed_tensor(x, m, 0.1, 8, mode)
        c2[name] = {"embed_kernel_us": round(kernel_us(lambda: W.embed_tensor(x, m, 0.1, 8, mode, out=o)), 1),
                    "extract_kernel_us": round(kernel_us(lambda: W.extract_tensor(o, x, 0.1, 8, mode)), 1),
                    "embed_ms_e2e_pil": round(median_ms(lambda: W.embed_watermark(img4, png4, True, s), 10), 2),
                    "extract_ms_e2e_pil": round(median_ms(lambda: W.extract_watermark(img4, img4, s), 10), 2)}
        ext4 = W.extract_tensor(o, x, 0.1, 8, mode).cpu().numpy()
        c2[name]["payload_ok"] = Q.decode_map(ext4) == Q.encrypt("Test" * 10)
    c2["pil_only_ms"] = {"convert_and_resize_watermark": round(median_ms(
        lambda: np.array(W.resize_watermark(png4, 270, 480, True)), 10, sync=False), 2)}
    res["config2_4k_single"] = c2

    # ---- config 4: extract of watermarked 1080p batch, payload check on every natural image
    n = 64
    imgs = torch.empty((n, bench.H, bench.W, 3), dtype=torch.uint8, device="cuda")
    bench.fill_images_device(imgs, 0, 17)
    payload = Q.encrypt("Test" * 10)
    wm1080 = np.array(W.resize_watermark(Q.qr_png(payload), 135, 240, True))
    wmd = torch.from_numpy(wm1080).cuda()
    c4 = {}
    for mode, name in ((1, "fast"), (0, "faithful")):
        out = W.embed_tensor(imgs, wmd, 0.1, 8, mode)
        ext = W.extract_tensor(out, imgs, 0.1, 8, mode).cpu().numpy()
        ok = {k: 0 for k in ("natural", "random", "regions")}
        tot = {k: 0 for k in ok}
        for k in range(n):
            kind = bench.image_kind(k)
            tot[kind] += 1
            ok[kind] += int(Q.decode_map(ext[k]) == payload)
        # reference extractor (oracle) on 4 of the GPU's images: same payload bytes
        cross = [Q.decode_map(O.extract_array(out[k].cpu().numpy(), imgs[k].cpu().numpy())) == payload for k in (0, 1, 4, 5)]
        c4[name] = {"payload_byte_exact": ok, "of": tot, "oracle_extract_of_gpu_embed_ok": cross}
    c4["note"] = ("random / saturated-region images clip at white or have sigma0-ties, so the QR is not expected "
                  "to survive there in the reference either; natural images must all decode")
    res["config4_extract_batch"] = c4

    # ---- config 5: batched 8x8 SVD sweep
    Y = O.rgb_to_ycbcr(natural(2048, 4096, 5))[:, :, 0]
    D_all = torch.from_numpy(O.dct_blocks(O.to_blocks(Y)).reshape(-1, 8, 8)).cuda()
    c5 = []
    for nb in (1024, 4096, 32400, 129600, 518400, 1000000):
        reps = (nb + D_all.shape[0] - 1) // D_all.shape[0]
        D = D_all.repeat(reps, 1, 1)[:nb].contiguous()
        us_v = kernel_us(lambda: W.svd8x8(D, vectors=False), 30)
        us_f = kernel_us(lambda: W.svd8x8(D, vectors=True), 30)
        (_, _, _), sw = W.svd8x8(D[: min(nb, 32400)], vectors=True, return_sweeps=True)
        c5.append({"blocks": nb, "values_only_us": round(us_v, 1), "full_us": round(us_f, 1),
                   "values_only_Mblocks_per_s": round(nb / us_v, 2), "full_Mblocks_per_s": round(nb / us_f, 2),
                   "values_only_GBps_algorithmic": round(nb * 288 / us_v / 1e3, 1),
                   "full_GBps_algorithmic": round(nb * 800 / us_f / 1e3, 1),
                   "sweeps_mean": round(float(sw.float().mean()), 2), "sweeps_max": int(sw.max())})
    res["config5_svd_sweep"] = c5
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
