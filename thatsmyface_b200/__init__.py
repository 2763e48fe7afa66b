"""thatsmyface_b200 - B200-native (sm_100a) DCT+SVD watermark embed/extract.

A drop-in for the hot path of Rigelyon/ThatsMyFace ``modules/watermarking.py``:

    from thatsmyface_b200.watermarking import embed_watermark, extract_watermark

Everything else of the product (face recognition, fuzzy extractor, AES, QR
generation/decoding, the Streamlit UI) stays on the Python host, unchanged.
"""
from .constants import ALPHA, BLOCK_SIZE, MODE_FAITHFUL, MODE_FAST, MODE_LITERAL  # noqa: F401

__version__ = "0.1.0"
