"""ctypes binding of libtmfwm.so (include/tmf_wm.h).

The library is the product: if it is missing or cannot be loaded this module
raises - there is no Python/NumPy fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIBPATH

_vp, _i, _i64, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double

# name -> (restype, argtypes); exactly the symbols include/tmf_wm.h declares
PROTOTYPES = {
    "tmf_version": (_i, []),
    "tmf_last_error": (C.c_char_p, []),
    "tmf_device_count": (_i, []),
    "tmf_last_fast_path": (_i, []),
    "tmf_embed_rgb8": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _i, _d, _i, _i, _vp]),
    "tmf_extract_rgb8": (_i, [_vp, _vp, _vp, _i, _i, _i, _sz, _d, _i, _i, _vp]),
    "tmf_sigma0_rgb8": (_i, [_vp, _vp, _i, _i, _i, _sz, _i, _i, _vp]),
    "tmf_svd8x8_f32": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i, _vp]),
    "tmf_dct8x8_f32": (_i, [_vp, _vp, _i64, _i, _vp]),
    "tmf_rgb8_to_ycbcr_f32": (_i, [_vp, _vp, _i64, _vp]),
    "tmf_ycbcr_f32_to_rgb8": (_i, [_vp, _vp, _i64, _vp]),
    "tmf_rgbx8_to_rgb8": (_i, [_vp, _vp, _i64, _vp]),
    "tmf_rgb8_to_rgbx8": (_i, [_vp, _vp, _i64, _i, _vp]),
    "tmf_wm_map_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "tmf_wm_map_axis_table": (_i, [_i, _i, _vp, _vp, _vp, _sz]),
    "tmf_wm_map_l8": (_i, [_vp, _i, _i, _i, _sz, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "tmf_ctx_create": (_i, [C.POINTER(_vp), _i, _sz, _i]),
    "tmf_ctx_destroy": (_i, [_vp]),
    "tmf_ctx_embed_host_async": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _d, _i, _i]),
    "tmf_ctx_extract_host_async": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _i, _i]),
    "tmf_ctx_synchronize": (_i, [_vp]),
    "tmf_ctx_stats": (_i, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _i]),
    "tmf_pin_host": (_i, [_vp, _sz]),
    "tmf_unpin_host": (_i, [_vp]),
}

ERR_BAD_ARG, ERR_UNSUPPORTED_BLOCK, ERR_CUDA = -1, -2, -3

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("TMF_LIBPATH", LIBPATH)   # tuning sweeps load alternative builds of the same source
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m thatsmyface_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback for the watermark path."
            )
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().tmf_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map the C ABI's error codes onto the exceptions the reference's pages
    catch (they display ``str(e)``, embed_watermark_page.py:678-681)."""
    if rc == 0:
        return
    msg = last_error()
    if rc in (ERR_BAD_ARG, ERR_UNSUPPORTED_BLOCK):
        raise ValueError(msg)
    raise RuntimeError(msg)
