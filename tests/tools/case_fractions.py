"""Differing-sample fractions of the seeded cases of tests/test_gpu_parity.py::test_embed_extract_vs_oracle,
per content kind and mode (run on the GPU box; the thresholds in the test are set from this output)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import test_gpu_parity as T
from oracle import wm_oracle as O

res = {}
for kind, shape in T.ORACLE_CASES:
    rgb, wm = T.oracle_case(kind, shape)
    taps = {}
    ref = O.embed_array(rgb, wm, taps=taps)
    for mode in T.MODES:
        out = T.gpu_embed(rgb, wm, mode=mode)
        frac = T.assert_pixels(out, ref, taps["S"], f"{kind}{shape} mode {mode}")
        res[f"{kind}{shape} mode {mode}"] = round(frac, 6)
print(json.dumps(res, indent=1))
