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
VARIANTS = {          # the last sweeps of round 1 (tables 14-15 of r01_sweep_variants.txt); edit for the next one
    "base": {},
    "c6": {"TMF_EMBED_MIN_CTAS": 6},
    "c4": {"TMF_EMBED_MIN_CTAS": 4},
    "biasq": {"TMF_QUANT_DENORM": 0},
    "stash0": {"TMF_EMBED_STASH": 0},
    "rp1": {"TMF_EMBED_REPREFETCH": 1},
    "rowptr": {"TMF_EMBED_ROWPTR": 1},
    "p2u2": {"TMF_ROW_UNROLL_P2": 2},
    "x5": {"TMF_FAST_MIN_CTAS": 5},
    "t128": {"TMF_EMBED_THREADS": 128, "TMF_EXTRACT_THREADS": 128},
    "persist5": {"TMF_EMBED_PERSIST": 5},
}


def build():
    from thatsmyface_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    items = list(VARIANTS.items())
    for s in range(0, len(items), 6):
        procs = []
        for name, defs in items[s:s + 6]:
            out = os.path.join(VDIR, f"libtmfwm_{name}.so")
            cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-Xptxas", "-v"] + [f"-D{k}={v}" for k, v in defs.items()] + \
                [os.path.join(b.CSRC, src) for src in b.SOURCES] + ["-o", out, "-lcudart"]
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for name, p in procs:
            o = p.communicate()[0]
            regs = []
            lines = o.splitlines()
            for i, l in enumerate(lines):
                if "Compiling entry function" in l and ("embed_fastILi8" in l or "extract_fastILi8" in l or "_tma" in l):
                    regs.append(" ".join(x.strip() for x in lines[i + 2:i + 4]))
            print(name, "ok" if p.returncode == 0 else "FAILED\n" + o[-2000:], "|", " || ".join(regs))


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
