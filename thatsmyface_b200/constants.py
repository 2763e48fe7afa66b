"""Defaults the watermark path reads.

The numeric values are the ones the reference ships in ``modules/constants.py``
(lines 2-8); they are part of the drop-in contract (a user who never opens the
advanced settings must get the same block size and strength).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class WatermarkDefaults:
    block_size: int = 8          # modules/constants.py:7  (BLOCK_SIZE)
    alpha: float = 0.1           # modules/constants.py:8  (ALPHA)
    qrcode_size: int = 1000      # modules/constants.py:4  (QRCODE_SIZE), side of the QR image in pixels
    max_images: int = 30         # modules/constants.py:2  (MAX_IMAGES)
    max_watermark_characters: int = 100   # modules/constants.py:3


DEFAULTS = WatermarkDefaults()

# names the reference's modules import
BLOCK_SIZE = DEFAULTS.block_size
ALPHA = DEFAULTS.alpha
QRCODE_SIZE = DEFAULTS.qrcode_size
MAX_IMAGES = DEFAULTS.max_images
MAX_WATERMARK_CHARACTERS = DEFAULTS.max_watermark_characters

# `mode` of the fused kernels (include/tmf_wm.h)
MODE_FAITHFUL = 0   # DCT -> one-sided Jacobi SVD -> IDCT, bit-exact colour
MODE_FAST = 1       # spatial top-triplet + rank-1 update (default); no DCT / SVD executed
MODE_LITERAL = 2    # block size 8: faithful with the literal U diag(S') V^T product (V accumulated)
