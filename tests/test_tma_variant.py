"""The TMA-tiled persistent embed kernel (k_embed_tile, the default for 16-byte aligned batches whose
rows hold a multiple of 16 blocks) against the per-thread kernel on the same pixels: the outputs
must be identical (tests/tools/ab_tma.py: aligned buffers take the tile kernel, the same data at an
8-byte offset cannot).  Extract has one kernel; the tool checks it is insensitive to the offset too."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1080, 1920),     # 15 boxes per block-row: every other tile wraps to the next row
                                   (3, 264, 512),       # 4 boxes per row
                                   (1, 8, 128),         # one box: half a tile
                                   (5, 24, 384),        # 3 boxes per row, 9 per image: tiles straddle images
                                   (37, 2160, 3840)])   # more tiles than resident warps: the persistent loop
def test_tma_kernel_equals_per_thread_kernel(shape):
    n, h, w = shape
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ab_tma.py"), str(n), str(h), str(w)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    counts = [int(m) for m in re.findall(r"mismatching samples (\d+) of", r.stdout)]
    assert len(counts) == 3 and counts == [0, 0, 0], r.stdout
    m = re.search(r"extract paths \([01], 0\): mismatching levels (\d+) of", r.stdout)
    assert m and int(m.group(1)) == 0, r.stdout
