"""Golden vectors for the watermark-map row, from the LIVE reference.

TEST INFRASTRUCTURE ONLY; build container only (imports /root/reference through
``oracle/live_reference.py``).  Writes ``tests/golden/wm_map_cases.npz``: a few mode-"L"
watermark sources and, for each (source, target, preserve_ratio), the map the reference's own
``resize_watermark`` (modules/watermarking.py:86-132) returns.  Also checks the restatement in
``oracle/pil_lanczos.py`` against every one of them before writing.

    python -m oracle.make_golden_wm_map
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import PIL
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import live_reference, pil_lanczos  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def sources():
    import qr_util  # tests/qr_util.py (cv2 stand-in for the reference's QR generator)

    rng = np.random.default_rng(11)
    qr = np.array(Image.open(__import__("io").BytesIO(qr_util.qr_png(qr_util.encrypt("golden map"), 1000))).convert("L"))
    yy, xx = np.mgrid[0:61, 0:173]
    return {
        "qr1000": qr,                                                         # the page's 1000x1000 QR
        "noise_61x173": rng.integers(0, 256, (61, 173), dtype=np.uint8),
        "ramp_61x173": ((xx * 3 + yy * 5) % 256).astype(np.uint8),
        "tiny_3x5": rng.integers(0, 256, (3, 5), dtype=np.uint8),
    }


# (source, target_h, target_w, preserve_ratio)
CASES = [
    ("qr1000", 135, 240, True), ("qr1000", 135, 240, False), ("qr1000", 64, 64, True), ("qr1000", 270, 480, True),
    ("qr1000", 1000, 1000, False), ("qr1000", 90, 51, True),
    ("noise_61x173", 16, 25, True), ("noise_61x173", 16, 25, False), ("noise_61x173", 61, 40, False),
    ("noise_61x173", 30, 173, False), ("noise_61x173", 200, 300, True), ("ramp_61x173", 135, 240, True),
    ("tiny_3x5", 17, 9, False), ("tiny_3x5", 2, 2, False),
]


def main():
    R = live_reference.load()
    src = sources()
    out = {f"src_{k}": v for k, v in src.items()}
    meta = []
    for i, (name, th, tw, pr) in enumerate(CASES):
        ref = np.array(R.resize_watermark(Image.fromarray(src[name], "L"), th, tw, pr))
        mine = pil_lanczos.watermark_map_l8(src[name], th, tw, pr)
        assert ref.shape == (th, tw) and np.array_equal(ref, mine), (name, th, tw, pr)
        out[f"map_{i}"] = ref
        meta.append(dict(source=name, target_h=th, target_w=tw, preserve_ratio=bool(pr),
                         sha256=hashlib.sha256(ref.tobytes()).hexdigest()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "wm_map_cases.npz"), **out)
    mpath = os.path.join(GOLDEN_DIR, "MANIFEST.json")
    manifest = json.load(open(mpath))
    manifest["wm_map_cases"] = dict(pillow=PIL.__version__, restatement_identical=True, cases=meta)
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)
    print(f"{len(meta)} map cases written; restatement identical on all")


if __name__ == "__main__":
    main()
