"""ctypes wrapper over tests/hostsim/_hostsim.so - the kernel arithmetic of
thatsmyface_b200/csrc/tmf_math.cuh compiled for the host (TEST HARNESS ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
SO = os.path.join(HERE, "_hostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "..", "thatsmyface_b200", "csrc")
HDRS = [os.path.join(CSRC, n) for n in ("tmf_math.cuh", "tmf_fast.cuh", "tmf_resize.cuh")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        stale = (not os.path.exists(SO)) or any(os.path.getmtime(p) > os.path.getmtime(SO) for p in [SRC] + HDRS)
        if stale:
            subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO, "-lm"],
                           check=True)
        _lib = C.CDLL(SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def embed(rgb, wm, alpha=0.1, mode=0):
    rgb = np.ascontiguousarray(rgb)
    wm = np.ascontiguousarray(wm)
    h, w = rgb.shape[:2]
    nb = max(1, (h // 8) * (w // 8))
    out = np.empty_like(rgb)
    sig = np.zeros(nb, np.float32)
    sw = np.zeros(nb, np.int32)
    # mode 0: faithful with the literal U S' Vt product (the library's TMF_MODE_LITERAL); 1: fast;
    # 3: faithful as the library's default TMF_MODE_FAITHFUL computes it (no V, rank-1 reconstruction)
    fn = {0: lib().hostsim_embed, 1: lib().hostsim_embed_fast, 3: lib().hostsim_embed_rank1}[mode]
    fn(_p(rgb), _p(out), h, w, _p(wm), C.c_double(alpha), _p(sig), _p(sw))
    return out, sig[: (h // 8) * (w // 8)].reshape(h // 8, w // 8), sw


def extract(a, b, alpha=0.1, mode=0):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    h, w = a.shape[:2]
    out = np.zeros((h // 8, w // 8), np.uint8)
    fn = lib().hostsim_extract_fast if mode == 1 else lib().hostsim_extract
    fn(_p(a), _p(b), _p(out), h, w, C.c_double(alpha))
    return out


def svd(blocks):
    blocks = np.ascontiguousarray(blocks, np.float32).reshape(-1, 8, 8)
    n = len(blocks)
    AV, V, S = np.empty_like(blocks), np.empty_like(blocks), np.empty((n, 8), np.float32)
    sw = np.zeros(n, np.int32)
    lib().hostsim_svd(_p(blocks), C.c_int64(n), _p(AV), _p(V), _p(S), _p(sw))
    return AV, V, S, sw


def top_column(blocks):
    blocks = np.ascontiguousarray(blocks, np.float32).reshape(-1, 8, 8)
    n = len(blocks)
    s0, u0, sw = np.empty(n, np.float32), np.empty((n, 8), np.float32), np.zeros(n, np.int32)
    lib().hostsim_top_column(_p(blocks), C.c_int64(n), _p(s0), _p(u0), _p(sw))
    return s0, u0, sw


def dct(blocks, inverse=False):
    blocks = np.ascontiguousarray(blocks, np.float32).reshape(-1, 8, 8)
    out = np.empty_like(blocks)
    lib().hostsim_dct(_p(blocks), _p(out), C.c_int64(len(blocks)), 1 if inverse else 0)
    return out


def rgb2ycc(rgb):
    rgb = np.ascontiguousarray(rgb)
    out = np.empty(rgb.shape, np.float32)
    lib().hostsim_rgb2ycc(_p(rgb), _p(out), C.c_int64(rgb.shape[0] * rgb.shape[1]))
    return out


def embed_n(rgb, wm, alpha, bs):
    rgb, wm = np.ascontiguousarray(rgb), np.ascontiguousarray(wm)
    h, w = rgb.shape[:2]
    out = np.empty_like(rgb)
    assert lib().hostsim_embed_n(_p(rgb), _p(out), h, w, _p(wm), C.c_double(alpha), bs) == 0
    return out


def extract_n(a, b, alpha, bs):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    h, w = a.shape[:2]
    out = np.zeros((h // bs, w // bs), np.uint8)
    assert lib().hostsim_extract_n(_p(a), _p(b), _p(out), h, w, C.c_double(alpha), bs) == 0
    return out


def wm_map_l8(src, target_h, target_w, preserve_ratio=False):
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty((target_h, target_w), np.uint8)
    rc = lib().hostsim_wm_map_l8(_p(src), src.shape[0], src.shape[1], _p(out), target_h, target_w,
                                 1 if preserve_ratio else 0)
    if rc:
        raise ValueError("height and width must be > 0")
    return out
