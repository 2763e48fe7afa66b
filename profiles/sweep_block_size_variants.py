"""Tuning variants of the generic-N fast kernels (csrc/fast_n_kernels.cu), timed with block_size_sweep.py.

    python profiles/sweep_block_size_variants.py build      (build container)
    python profiles/sweep_block_size_variants.py run        (GPU box)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "thatsmyface_b200", "lib", "variants")
VARIANTS = {          # earlier tables: r02_sweep_j_block_sizes.txt (rows ahead, CTAs per SM), r02_sweep_l (packed Gram)
    "fn_default": {},
    "fn_14c2": {"TMF_FASTN_CTAS_14": 2},
    "fn_16c3": {"TMF_FASTN_CTAS_16": 3},
    "fn_12c2": {"TMF_FASTN_CTAS_12": 2},
}
SIZES = ["4", "6", "10", "12", "14", "16"]

if sys.argv[1] == "build":
    from thatsmyface_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        b.build(defines=[f"{k}={v}" for k, v in defs.items()], out=os.path.join(VDIR, f"libtmfwm_{name}.so"))
        print(name, "ok", flush=True)
else:
    for name in VARIANTS:
        env = dict(os.environ, TMF_LIBPATH=os.path.join(VDIR, f"libtmfwm_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "block_size_sweep.py")] + SIZES, env=env,
                           capture_output=True, text=True, timeout=600)
        print(f"== {name} {VARIANTS[name]}", flush=True)
        print(r.stdout.strip() if r.returncode == 0 else "FAILED\n" + r.stderr[-1500:], flush=True)
