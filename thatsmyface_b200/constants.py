"""Defaults of the watermark path - the values of the reference's
``modules/constants.py:2-9`` that the hot path reads."""

MAX_IMAGES = 30                  # constants.py:2
MAX_WATERMARK_CHARACTERS = 100   # constants.py:3
QRCODE_SIZE = 1000               # constants.py:4
BLOCK_SIZE = 8                   # constants.py:7
ALPHA = 0.1                      # constants.py:8

# fused-kernel modes (include/tmf_wm.h)
MODE_FAITHFUL = 0
MODE_FAST = 1
