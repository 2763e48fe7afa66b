"""In-tree build of libtmfwm.so (nvcc, sm_100a).  Used by ``__graft_entry__.build()``
and by ``python -m thatsmyface_b200.build``.  The .so stays in the tree
(``thatsmyface_b200/lib/``) so it travels to the GPU box with the snapshot.

One object per translation unit, compiled in parallel, then linked; objects are cached under
``thatsmyface_b200/lib/obj/`` keyed by source + flags, so touching one kernel family rebuilds one
file.  ``build(defines=[...], out=...)`` makes an alternative build of the same source with
``-D`` overrides of csrc/tmf_tunables.h (profiles/sweep_variants.py loads those through
TMF_LIBPATH)."""
from __future__ import annotations

import glob
import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBPATH = os.path.join(LIBDIR, "libtmfwm.so")
SOURCES = ["api.cu", "fast_kernels.cu", "fast_n_kernels.cu", "faithful_kernels.cu", "taps.cu", "wm_map.cu", "ctx.cu"]

# Never add --use_fast_math / -ftz=true here: the fast embed kernels' quantiser floors onto
# SUBNORMAL floats (the integer level is the float's bit pattern, csrc/tmf_rowmath.cuh
# embed_row_fast2); flushing them to zero would zero every output pixel.
# -ffp-contract=off: the host code builds Pillow's float64 Lanczos weight tables (tmf_resize.cuh),
# whose 22-bit fixed-point rounding must match Pillow bit for bit on any host (gcc contracts
# a*b+c by default on FMA-baseline targets); tests/hostsim is built with the same flag.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtmfwm.so cannot be built")


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(ROOT, "include", "*.h")))


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES] + _headers()


def is_stale() -> bool:
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def _closure(path: str, seen: list | None = None) -> list:
    """`path` and the project headers it includes, transitively (csrc/ and include/)."""
    seen = [] if seen is None else seen
    if path in seen or not os.path.isfile(path):
        return seen
    seen.append(path)
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
        for base in (CSRC, os.path.join(ROOT, "include")):
            _closure(os.path.join(base, inc), seen)
    return seen


def _key(src: str, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for p in _closure(os.path.join(CSRC, src)):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()[:20]


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    out = out or LIBPATH
    if not force and not defines and out == LIBPATH and not is_stale():
        return LIBPATH
    os.makedirs(os.path.dirname(out), exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    # a -D override of a tunable matters only to the translation units that use it, in their own text or in a
    # header they include (the tunables header itself merely supplies defaults): the others keep their objects
    def flags_for(src: str):
        text = "".join(open(p).read() for p in _closure(os.path.join(CSRC, src)) if not p.endswith("tmf_tunables.h"))
        used = [d for d in defines if d.split("=")[0] in text]
        return NVCC_FLAGS + [f"-D{d}" for d in used] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src: str) -> str:
        flags = flags_for(src)
        obj = os.path.join(OBJDIR, f"{os.path.splitext(src)[0]}.{_key(src, flags)}.o")
        if os.path.exists(obj) and not verbose and not force:
            return obj
        res = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    # keep the cache small: drop objects no current source/flag combination refers to, oldest first
    cached = sorted(glob.glob(os.path.join(OBJDIR, "*.o")), key=os.path.getmtime)
    for p in cached[:-64]:
        try:
            os.remove(p)
        except OSError:
            pass
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outp = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outp))
