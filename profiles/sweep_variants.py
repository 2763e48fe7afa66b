"""Build tuning variants of libtmfwm.so (same source, different -D tunables) and
time the fast kernels with each.  `build` runs in the build container (nvcc, no
GPU); `run` runs on the B200 box and prints one line per variant.

    python profiles/sweep_variants.py build
    python profiles/sweep_variants.py run [images]
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "thatsmyface_b200", "lib", "variants")
def tile(w, c, **kw):
    """Shape of the TMA-tiled embed kernel: warps per CTA (a multiple of 4), CTAs per SM."""
    d = {"TMF_TILE_WARPS": w, "TMF_TILE_CTAS_PER_SM": c}
    d.update(kw)
    return d


VARIANTS = {          # the knobs that are left (csrc/tmf_tunables.h); earlier sweeps of the round: profiles/r02_sweep_*.txt
    "base": {},
    "e_8x2": tile(8, 2),
    "e_12x1": tile(12, 1),
    "e_p2u2": {"TMF_ROW_UNROLL_P2": 2},
    "e_p2u8": {"TMF_ROW_UNROLL_P2": 8},
    "pt_notile": {},  # per-thread kernels (run with TMF_NO_TILE=1)
}


def build():
    from thatsmyface_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"libtmfwm_{name}.so")
        try:
            b.build(defines=[f"{k}={v}" for k, v in defs.items()], out=out)
            print(name, "ok", flush=True)
        except RuntimeError as e:
            print(name, "FAILED", str(e)[-1500:], flush=True)


def run_one(images):
    """Child process: TMF_LIBPATH selects the variant."""
    import torch
    import bench
    from thatsmyface_b200 import watermarking as W
    dev = torch.device("cuda", 0)
    imgs = torch.empty((images, bench.H, bench.W, 3), dtype=torch.uint8, device=dev)
    bench.fill_images_device(imgs, 0, 17)
    wm = torch.from_numpy(bench.make_wm_map()).to(dev)
    out = torch.empty_like(imgs)
    ext = torch.empty((images, bench.H // 8, bench.W // 8), dtype=torch.uint8, device=dev)
    res = {}
    mode = int(os.environ.get("TMF_SWEEP_MODE", "1"))
    for name, fn in (("embed", lambda: W.embed_tensor(imgs, wm, 0.1, 8, mode, out=out)),
                     ("extract", lambda: W.extract_tensor(out, imgs, 0.1, 8, mode, out=ext))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(images * bench.PX * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e6)
    print(json.dumps(res))


def run(images):
    names = sorted(f[len("libtmfwm_"):-3] for f in os.listdir(VDIR) if f.endswith(".so"))
    for name in names:
        env = dict(os.environ, TMF_LIBPATH=os.path.join(VDIR, f"libtmfwm_{name}.so"))
        if name.endswith("notile"):
            env["TMF_NO_TILE"] = "1"
        r = subprocess.run([sys.executable, __file__, "one", str(images)], env=env, capture_output=True, text=True)
        ok = r.returncode == 0 and r.stdout.strip()
        print(name, r.stdout.strip().splitlines()[-1] if ok else "FAILED " + r.stderr[-300:], flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]))
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 256)
