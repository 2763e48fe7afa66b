"""(Under tests/: the oracle is the checker.)  Randomised parity campaign on the GPU box: many small images of mixed content,
every mode (fast, faithful; literal too at block size 8), several alphas and block sizes, CUDA path (through the C ABI) vs
the oracle.  Prints one JSON summary with the differing fraction split by content kind AND mode.
python tests/tools/parity_campaign.py [images_per_kind]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import wm_oracle as O            # checker only
from thatsmyface_b200 import watermarking as W


def make(kind, h, w, rng):
    if kind == "random":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "natural":
        y, x = np.mgrid[0:h, 0:w].astype(np.float64)
        f1, f2 = rng.uniform(20, 120, 2)
        img = (rng.uniform(60, 180) + rng.uniform(10, 70) * np.sin(x / f1) * np.cos(y / f2))[..., None] \
            + rng.uniform(-20, 20, 3) + rng.normal(0, rng.uniform(0, 12), (h, w, 3))
        return np.clip(img, 0, 255).astype(np.uint8)
    if kind == "gray":
        g = make("natural", h, w, rng)[:, :, :1]
        return np.repeat(g, 3, axis=2)
    if kind == "dark":       # near-black with sparse bright pixels: near-tied singular values
        img = (rng.random((h, w, 3)) < 0.02).astype(np.uint8) * rng.integers(1, 256, (h, w, 3), dtype=np.uint8)
        return img.astype(np.uint8)
    if kind == "sat":        # mostly saturated
        return np.clip(rng.normal(245, 15, (h, w, 3)), 0, 255).astype(np.uint8)
    if kind == "edges":      # blocks straddling hard edges
        img = np.zeros((h, w, 3), np.uint8)
        img[:, :: rng.integers(3, 17)] = 255
        img[:: rng.integers(3, 17)] = rng.integers(0, 256, 3, dtype=np.uint8)
        return img
    raise ValueError(kind)


def main(per_kind):
    rng = np.random.default_rng(2026)
    kinds = ("random", "natural", "gray", "dark", "sat", "edges")
    res = {"images": 0, "max_pixel_diff_outside_ties": 0, "max_extract_diff": 0, "bit_mismatches_decided": 0,
           "tie_blocks_excluded": 0, "blocks": 0, "worst": None, "by_kind": {}, "by_kind_and_mode": {}}
    mode_names = {0: "faithful", 1: "fast", 2: "literal"}
    for kind in kinds:
        agg = {"n": 0, "frac_px_differing": [], "max_px": 0, "max_ext": 0}
        per_mode = {m: {"pixels": 0, "pixels_differing": 0, "samples": 0, "samples_differing": 0, "ext": 0, "ext_differing": 0}
                    for m in mode_names}
        for k in range(per_kind):
            bs = 8 if k % 4 else int(rng.choice([4, 6, 10, 12, 14, 16]))
            h, w = int(rng.integers(2, 20)) * bs + int(rng.integers(0, bs)), int(rng.integers(2, 24)) * bs + int(rng.integers(0, bs))
            alpha = float(rng.choice([0.1, 0.1, 0.2, 0.5, 1.0]))
            img = make(kind, h, w, rng)
            wm = rng.integers(0, 256, (h // bs, w // bs), dtype=np.uint8)
            wm[rng.random(wm.shape) < 0.3] = 0
            wm[rng.random(wm.shape) < 0.3] = 255
            taps = {}
            ref = O.embed_array(img, wm, alpha, bs, taps=taps)
            ref_ext = O.extract_array(ref, img, alpha, bs)
            S = taps.get("S")
            tie = ((S[..., 0] - S[..., 1]) <= 1e-4 * np.maximum(S[..., 0], 1e-30)) & (S[..., 0] > 0)
            tie_px = np.zeros((h, w), bool)
            t = np.repeat(np.repeat(tie, bs, 0), bs, 1)
            tie_px[: t.shape[0], : t.shape[1]] = t
            x = torch.from_numpy(img).cuda()
            for mode in ((0, 1, 2) if bs == 8 else (0, 1)):
                out = W.embed_tensor(x, torch.from_numpy(wm).cuda(), alpha, bs, mode).cpu().numpy()
                d3 = np.abs(out.astype(int) - ref.astype(int))
                d3[tie_px] = 0
                d = d3.max(axis=2)
                ext = W.extract_tensor(torch.from_numpy(ref).cuda(), x, alpha, bs, mode).cpu().numpy()
                de = np.abs(ext.astype(int) - ref_ext.astype(int))
                decided = np.abs(ref_ext.astype(int) - 128) > 1
                bad_bits = int(((ext >= 128) != (ref_ext >= 128))[decided].sum())
                agg["max_px"] = max(agg["max_px"], int(d.max()))
                agg["max_ext"] = max(agg["max_ext"], int(de.max()) if de.size else 0)
                agg["frac_px_differing"].append(float((d > 0).mean()))
                pm = per_mode[mode]
                pm["pixels"] += d.size; pm["pixels_differing"] += int((d > 0).sum())
                pm["samples"] += d3.size; pm["samples_differing"] += int((d3 > 0).sum())
                pm["ext"] += de.size; pm["ext_differing"] += int((de > 0).sum())
                res["bit_mismatches_decided"] += bad_bits
                if d.max() > 1 or (de.size and de.max() > 1) or bad_bits:
                    res["worst"] = {"kind": kind, "k": k, "bs": bs, "alpha": alpha, "mode": mode, "shape": [h, w],
                                    "max_px": int(d.max()), "max_ext": int(de.max()), "bad_bits": bad_bits}
            agg["n"] += 1
            res["images"] += 1
            res["blocks"] += int(tie.size)
            res["tie_blocks_excluded"] += int(tie.sum())
        res["by_kind"][kind] = {"n": agg["n"], "max_pixel_diff": agg["max_px"], "max_extract_diff": agg["max_ext"],
                                "mean_fraction_of_pixels_differing_by_1": round(float(np.mean(agg["frac_px_differing"])), 6)}
        res["by_kind_and_mode"][kind] = {
            mode_names[m]: {"fraction_of_pixels_differing_by_1": round(v["pixels_differing"] / max(v["pixels"], 1), 6),
                            "fraction_of_samples_differing_by_1": round(v["samples_differing"] / max(v["samples"], 1), 6),
                            "fraction_of_extracted_levels_differing_by_1": round(v["ext_differing"] / max(v["ext"], 1), 6)}
            for m, v in per_mode.items()}
        res["max_pixel_diff_outside_ties"] = max(res["max_pixel_diff_outside_ties"], agg["max_px"])
        res["max_extract_diff"] = max(res["max_extract_diff"], agg["max_ext"])
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 60)
