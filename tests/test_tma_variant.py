"""The opt-in TMA-staged embed kernel (TMF_EMBED_TMA=1, k_embed_fast_tma) against the default
per-thread kernel on the same pixels: the outputs must be identical.  The library reads the
environment variable once, so the comparison runs in a child process (tests/tools/ab_tma.py:
16-byte aligned buffers take the TMA kernel, the same data at an 8-byte offset cannot)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1080, 1920),     # 240 blocks per row: every other warp straddles a block-row
                                   (3, 264, 512),       # 64 per row: single runs
                                   (2, 72, 272)])       # 34 per row: runs of every even length
def test_tma_kernel_equals_per_thread_kernel(shape):
    n, h, w = shape
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ab_tma.py"), str(n), str(h), str(w)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    counts = [int(m) for m in re.findall(r"mismatching samples (\d+) of", r.stdout)]
    assert len(counts) == 3 and counts == [0, 0, 0], r.stdout
