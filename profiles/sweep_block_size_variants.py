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
VARIANTS = {          # earlier tables: r02_sweep_j (rows ahead, CTAs per SM), r02_sweep_l (packed Gram), r02_sweep_m (CTA size)
    "fn_default": {},          # one-warp CTAs for 10 ... 16, 128 threads for 4 and 6
    "fn_t128": {"TMF_FASTN_THREADS": 128, "TMF_FASTN_CTAS_10": 4, "TMF_FASTN_CTAS_12": 3, "TMF_FASTN_CTAS_14": 3, "TMF_FASTN_CTAS_16": 2},
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
