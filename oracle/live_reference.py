"""Import the UNMODIFIED reference modules: from /root/reference in the build container, or from
the byte-for-byte copy that ``oracle/stage_ref.py`` left in ``oracle/_ref/`` (git-ignored, travels
to the GPU box with the snapshot).

TEST INFRASTRUCTURE ONLY: ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``
use it as the checker / the timed CPU baseline; the product never imports it.

Imports the reference needs that are absent from this image are stubbed, none of them on the
watermark path: ``streamlit`` (``modules/watermarking.py:5``, used only for ``st.session_state``
in ``get_watermark_settings``, :10-20) and ``deepface`` (``modules/utils.py:9``,
``modules/face_recognition.py:4``: face detection, never called by the fuzzy extractor).
"""
from __future__ import annotations

import os
import sys
import types

from . import stage_ref

REFERENCE_ROOT = "/root/reference"


def root():
    """Directory that holds the reference's ``modules`` package, or None."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "watermarking.py")):
        return REFERENCE_ROOT
    if stage_ref.staged():
        return stage_ref.REF_DIR
    return None


def available() -> bool:
    return root() is not None


def kind() -> str:
    r = root()
    return "none" if r is None else ("tree" if r == REFERENCE_ROOT else "staged copy")


def _prepare():
    r = root()
    if r is None:
        raise RuntimeError("reference not present: neither " + REFERENCE_ROOT + " nor oracle/_ref (run oracle/stage_ref.py "
                           "where the reference tree exists)")
    if "streamlit" not in sys.modules:
        st = types.ModuleType("streamlit")
        st.session_state = {}
        sys.modules["streamlit"] = st
    if "deepface" not in sys.modules:
        df = types.ModuleType("deepface")
        df.DeepFace = types.SimpleNamespace()
        sys.modules["deepface"] = df
    if r not in sys.path:
        sys.path.insert(0, r)


def load():
    """The reference's ``modules.watermarking`` module object."""
    _prepare()
    from modules import watermarking  # noqa: E402  (the reference's own file)

    return watermarking


def load_fuzzy():
    """The reference's ``modules.fuzzy_extractor`` (generate_key_with_helper / regenerate_key_from_helper)."""
    _prepare()
    from modules import fuzzy_extractor  # noqa: E402

    return fuzzy_extractor
