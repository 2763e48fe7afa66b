"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden [--exhaustive-colour]

For every case it runs the reference's own ``embed_watermark`` /
``extract_watermark`` (``modules/watermarking.py:135,224``) through PIL, checks
that ``oracle.wm_oracle`` reproduces the reference bit for bit, and stores the
inputs and the reference's outputs as small fixtures.  The GPU box has no
/root/reference, so the ``-m gpu`` tests read only these files.
"""
from __future__ import annotations

import argparse
import hashlib
import io
import json
import os
import sys
import time

import numpy as np
from PIL import Image

from oracle import live_reference, wm_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SETTINGS = {"block_size": 8, "alpha": 0.1}


def natural_like(h, w, seed):
    """SURVEY.md 8(d) config-1 generator: smooth field + per-channel offset + noise."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = 120 + 70 * np.sin(x / 97.0) * np.cos(y / 71.0)
    img = base[..., None] + np.array([10.0, 0.0, -10.0]) + rng.normal(0, 8, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def regions(h, w, seed):
    """Flat / black / saturated / gradient / sparse-dot regions: degenerate blocks."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 4] = 0
    img[h // 4: h // 2, : w // 2] = 255
    img[h // 4: h // 2, w // 2:] = 128
    img[h // 2: 3 * h // 4, :, :] = (np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)[None, :, None]
    # sparse bright dots on black (near-tied singular values live here)
    img[2, 3] = (200, 10, 10)
    img[5, 6] = (10, 180, 30)
    img[9, 9] = (255, 255, 255)
    return img


def cases():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    w = rng.integers(0, 256, (8, 8), dtype=np.uint8)
    yield "gv1_random64", a, w
    r = np.random.default_rng(11)
    yield "natural_96x120", natural_like(96, 120, 2), r.integers(0, 256, (12, 15), dtype=np.uint8)
    yield "natural_ragged_70x93", natural_like(70, 93, 3), (r.integers(0, 2, (8, 11)) * 255).astype(np.uint8)
    yield "regions_64x80", regions(64, 80, 4), (r.integers(0, 2, (8, 10)) * 255).astype(np.uint8)
    for name, v in (("flat_black16", 0), ("flat_gray16", 128), ("flat_white16", 255)):
        yield name, np.full((16, 16, 3), v, np.uint8), np.full((2, 2), 255, np.uint8)  # KAT-flat
    yield "zero_wm_32", r.integers(0, 256, (32, 32, 3), dtype=np.uint8), np.zeros((4, 4), np.uint8)  # KAT-zero-wm
    yield "tiny_8x8", r.integers(0, 256, (8, 8, 3), dtype=np.uint8), np.array([[255]], np.uint8)
    yield "subblock_5x7", r.integers(0, 256, (5, 7, 3), dtype=np.uint8), np.zeros((0, 0), np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exhaustive-colour", action="store_true",
                    help="also check rgb_to_ycbcr over all 2**24 RGB triplets (~1 min)")
    args = ap.parse_args()
    R = live_reference.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    report = {"settings": SETTINGS, "cases": {}, "numpy": np.__version__}
    import scipy, PIL
    report["scipy"], report["pillow"] = scipy.__version__, PIL.__version__

    for name, rgb, wm in cases():
        img = Image.fromarray(rgb)
        nbh, nbw = rgb.shape[0] // 8, rgb.shape[1] // 8
        t0 = time.time()
        if nbh and nbw:
            wm_img = Image.fromarray(wm)  # already (nbw, nbh): PIL resize is a copy
            ref_out = np.array(R.embed_watermark(img, wm_img, False, dict(SETTINGS)))
            ref_ext = np.array(R.extract_watermark(Image.fromarray(ref_out), img, dict(SETTINGS)))
        else:  # no whole block: the reference's loops do not run; colour round trip only
            ref_out = R.ycbcr_to_rgb(R.rgb_to_ycbcr(img))
            ref_ext = np.zeros((nbh, nbw), np.uint8)
        dt = time.time() - t0
        taps = {}
        ora_out = O.embed_array(rgb, wm, SETTINGS["alpha"], 8, taps=taps)
        ora_ext = O.extract_array(ref_out, rgb, SETTINGS["alpha"], 8)
        ok_out, ok_ext = bool((ora_out == ref_out).all()), bool((ora_ext == ref_ext).all())
        # loop-style oracle must agree too
        loop_out = O.embed_array(rgb, wm, SETTINGS["alpha"], 8, style="loop")
        ok_loop = bool((loop_out == ref_out).all())
        # stage taps straight from the reference's helpers
        ref_ycc = R.rgb_to_ycbcr(img)
        ok_y = bool((taps["Y"] == ref_ycc[:, :, 0]).all())
        save = dict(rgb=rgb, wm=wm, ref_out=ref_out, ref_ext=ref_ext, ref_ycc=ref_ycc.astype(np.float32))
        if nbh and nbw:
            d00 = R.apply_dct_to_block(ref_ycc[:8, :8, 0])
            s00 = np.linalg.svd(d00, full_matrices=True)[1]
            ok_d = bool((taps["D"][0, 0] == d00).all()) and bool((taps["S"][0, 0] == s00).all())
            save.update(ref_S=taps["S"], ref_D00=d00)
        else:
            ok_d = True
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **save)
        report["cases"][name] = dict(
            shape=list(rgb.shape), sha_out=sha(ref_out), sha_ext=sha(ref_ext), ref_seconds=round(dt, 3),
            oracle_out_identical=ok_out, oracle_ext_identical=ok_ext, oracle_loop_identical=ok_loop,
            oracle_Y_identical=ok_y, oracle_DS_identical=ok_d,
        )
        print(name, report["cases"][name])
        if not (ok_out and ok_ext and ok_loop and ok_y and ok_d):
            print("ORACLE MISMATCH in", name, file=sys.stderr)

    # other block sizes / alphas of the UI sliders (embed_watermark_page.py:324-350)
    for bs, alpha in ((4, 0.1), (6, 0.3), (12, 0.5), (16, 1.0)):
        r2 = np.random.default_rng(100 + bs)
        rgb = natural_like(5 * bs + 2, 6 * bs + 1, 30 + bs)
        wm = np.where(r2.random((5, 6)) < 0.4, 0, r2.integers(1, 256, (5, 6))).astype(np.uint8)
        st = {"block_size": bs, "alpha": alpha}
        ref_out = np.array(R.embed_watermark(Image.fromarray(rgb), Image.fromarray(wm), False, dict(st)))
        ref_ext = np.array(R.extract_watermark(Image.fromarray(ref_out), Image.fromarray(rgb), dict(st)))
        ok_out = bool((O.embed_array(rgb, wm, alpha, bs) == ref_out).all())
        ok_ext = bool((O.extract_array(ref_out, rgb, alpha, bs) == ref_ext).all())
        name = f"bs{bs}_alpha{alpha}"
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), rgb=rgb, wm=wm, ref_out=ref_out, ref_ext=ref_ext,
                            bs=np.int32(bs), alpha=np.float64(alpha))
        report["cases"][name] = dict(shape=list(rgb.shape), sha_out=sha(ref_out), sha_ext=sha(ref_ext),
                                     oracle_out_identical=ok_out, oracle_ext_identical=ok_ext)
        print(name, report["cases"][name])

    # PIL-level path with a PNG-bytes watermark and LANCZOS + white padding
    rng = np.random.default_rng(5)
    qr_like = np.kron((rng.integers(0, 2, (25, 25)) * 255).astype(np.uint8), np.ones((8, 8), np.uint8))
    buf = io.BytesIO()
    Image.fromarray(qr_like).save(buf, format="PNG")
    png = buf.getvalue()
    rgb = natural_like(128, 200, 6)
    for pr in (True, False):
        ref_out = np.array(R.embed_watermark(Image.fromarray(rgb), png, pr, dict(SETTINGS)))
        ora_out = np.array(O.embed_watermark(Image.fromarray(rgb), png, pr, dict(SETTINGS)))
        wm_map = np.array(R.resize_watermark(png, 16, 25, pr))
        ref_ext = np.array(R.extract_watermark(Image.fromarray(ref_out), Image.fromarray(rgb), dict(SETTINGS)))
        name = f"pil_png_preserve{int(pr)}"
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), rgb=rgb, png=np.frombuffer(png, np.uint8),
                            wm=wm_map, ref_out=ref_out, ref_ext=ref_ext)
        report["cases"][name] = dict(shape=list(rgb.shape), sha_out=sha(ref_out), sha_ext=sha(ref_ext),
                                     oracle_out_identical=bool((ora_out == ref_out).all()))
        print(name, report["cases"][name])

    # forward colour transform: sampled always, exhaustive on request
    if args.exhaustive_colour:
        g = np.arange(256, dtype=np.uint8)
        mism = 0
        t0 = time.time()
        for r in range(256):
            tile = np.empty((256, 256, 3), np.uint8)
            tile[..., 0] = r
            tile[..., 1] = g[:, None]
            tile[..., 2] = g[None, :]
            mism += int((R.rgb_to_ycbcr(tile) != O.rgb_to_ycbcr(tile)).sum())
        report["colour_forward_exhaustive"] = dict(triplets=1 << 24, mismatching_values=mism,
                                                   seconds=round(time.time() - t0, 1))
        print("exhaustive colour:", report["colour_forward_exhaustive"])
    rng = np.random.default_rng(7)
    ycc = rng.random((300, 300, 3), dtype=np.float32) * np.float32(1.2) - np.float32(0.1)
    report["colour_inverse_sampled"] = dict(
        values=int(ycc.size), mismatching_values=int((R.ycbcr_to_rgb(ycc) != O.ycbcr_to_rgb(ycc)).sum()))
    print("inverse colour:", report["colour_inverse_sampled"])

    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
