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
        return NVCC_FLAGS + [f"-D{d}" for d in used] + ["-Xptxas", "-v"]

    def compile_one(src: str) -> str:
        flags = flags_for(src)
        obj = os.path.join(OBJDIR, f"{os.path.splitext(src)[0]}.{_key(src, flags)}.o")
        log = obj[:-2] + ".ptxas.txt"        # registers / stack / spills of every kernel (resource_report)
        if os.path.exists(obj) and os.path.exists(log) and not force:
            if verbose:
                print(open(log).read())
            return obj
        res = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        with open(log, "w") as f:
            f.write(res.stderr)
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if out == LIBPATH:                       # which objects the default library was linked from
        with open(os.path.join(LIBDIR, "linked_objects.txt"), "w") as f:
            f.write("\n".join(objs) + "\n")
    # keep the cache small: drop objects no current source/flag combination refers to, oldest first
    cached = sorted(glob.glob(os.path.join(OBJDIR, "*.o")), key=os.path.getmtime)
    for p in cached[:-64]:
        for q in (p, p[:-2] + ".ptxas.txt"):
            try:
                os.remove(q)
            except OSError:
                pass
    return out


def resource_report() -> list:
    """(kernel, registers, stack bytes, spill-store bytes, spill-load bytes) of every kernel of the default library,
    from the ptxas -v logs kept beside its objects."""
    import subprocess as sp
    rows = []
    listing = os.path.join(LIBDIR, "linked_objects.txt")
    if not os.path.exists(listing):
        return rows                           # a library built elsewhere (or before the logs were kept)
    for obj in open(listing).read().split():
        log = obj[:-2] + ".ptxas.txt"
        if not os.path.exists(log):
            continue
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\s*\n.*?Used (\d+) registers", text):
            rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
    names = sp.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
    out_rows = []
    for r, n in zip(rows, names):
        n = re.sub(r"tmfi::\(anonymous namespace\)::", "", n)
        n = re.sub(r"\(.*$", "", n).replace("void ", "")
        out_rows.append((n,) + r[1:])
    return sorted(set(out_rows))


if __name__ == "__main__" and "--resources" in sys.argv:
    for row in resource_report():
        print("%-48s regs %3d  stack %5d  spill st %5d ld %5d" % row)
    sys.exit(0)

if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outp = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outp))
