"""Fixture for the "helper data" leg of BASELINE config 4: the reference's own
``generate_key_with_helper`` (modules/fuzzy_extractor.py:194-228) run on the synthetic 512-d
embedding ``default_rng(0).normal(size=512)`` with the tolerance the embed page passes
(ERROR_TOLERANCE / 100 = 0.6, embed_watermark_page.py:430-438).  ``generate`` draws its sketch seed
from ``os.urandom``, so the helper is not reproducible: it is committed, with the key it belongs
to, as ``tests/golden/helper_case.json``.  The extract side
(``regenerate_key_from_helper``, :230-275, called at extract_watermark_page.py:266) then runs from
the staged reference on any box and must give the same key back.

    python oracle/make_golden_helper.py          # build container only (needs /root/reference)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import live_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "helper_case.json")


def embedding(seed=0):
    return np.random.default_rng(seed).normal(size=512)


def main():
    F = live_reference.load_fuzzy()
    emb = embedding()
    key, helper = F.generate_key_with_helper(emb, 0.6)
    again = F.regenerate_key_from_helper(emb + np.random.default_rng(1).normal(size=512) * 0.02, helper)
    assert again == key and len(key) == 32, "the reference's own round trip must work before the fixture is written"
    helper = {k: (list(v) if isinstance(v, tuple) else v) for k, v in helper.items()}
    with open(OUT, "w") as f:
        json.dump({"embedding": "numpy.random.default_rng(0).normal(size=512)", "error_tolerance": 0.6,
                   "helper": helper, "key_hex": key.hex()}, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; key", key.hex()[:16], "...")


if __name__ == "__main__":
    main()
