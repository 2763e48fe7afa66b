"""Stage the UNMODIFIED reference files of the watermark path into ``oracle/_ref/`` so that they
travel to the GPU box (``/root/reference`` does not): ``modules/watermarking.py`` +
``constants.py`` (the hot path) and ``fuzzy_extractor.py`` with the three modules
it imports (the "helper data" leg of BASELINE config 4).  The files are copied byte for byte -
never edited - and ``oracle/_ref/`` is git-ignored: the reference's sources are not part of this
repository's history.  ``STAGED.json`` records their SHA-256.

TEST INFRASTRUCTURE ONLY: used by ``oracle/live_reference.py`` (tests, ``smoke()``, the CPU legs of
``bench.py``); the product never imports it.  Run by ``__graft_entry__.build()``; a no-op where the
reference tree is absent (the GPU box uses the copy that came with the snapshot).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REFERENCE_ROOT = "/root/reference"
REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["modules/__init__.py", "modules/constants.py", "modules/watermarking.py", "modules/fuzzy_extractor.py",
         "modules/utils.py", "modules/face_recognition.py"]


def staged() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "modules", "watermarking.py"))


def stage(verbose: bool = False) -> bool:
    """Copy the files; returns True if ``oracle/_ref`` is usable afterwards."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "watermarking.py")):
        return staged()
    sums = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            sums[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF_DIR, "STAGED.json"), "w") as f:
        json.dump({"from": REFERENCE_ROOT, "sha256": sums}, f, indent=1, sort_keys=True)
    if verbose:
        print("staged", len(FILES), "reference files into", REF_DIR)
    return True


if __name__ == "__main__":
    print("oracle/_ref usable:", stage(verbose=True))
