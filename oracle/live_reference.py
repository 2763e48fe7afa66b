"""Import the UNMODIFIED reference module from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing that runs there may import this file; it is used by
``oracle/make_golden.py`` (fixture generation) and by CPU tests that skip when
the tree is absent.

The only missing import of ``modules/watermarking.py`` is ``streamlit`` (line 5),
used solely for ``st.session_state`` in ``get_watermark_settings`` (lines 10-20);
a stub module with an empty ``session_state`` is installed before the import.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "watermarking.py"))


def load():
    """Return the reference's ``modules.watermarking`` module object."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    if "streamlit" not in sys.modules:
        st = types.ModuleType("streamlit")
        st.session_state = {}
        sys.modules["streamlit"] = st
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from modules import watermarking  # noqa: E402  (the reference's own file)

    return watermarking
