"""QR + AES helpers for the end-to-end payload tests.  The reference uses the
`qrcode`, `pyzbar` and `pycryptodome` packages (modules/qrcode_generator.py:10-76,
modules/encryption.py:8-68), none of which is installed here; OpenCV's QR
encoder/decoder (ECC level H) and `cryptography` (AES-256-CBC + PKCS7) stand in.
These stay on the host in the product too - they only produce the watermark
image and read the extracted one."""
import base64
import io

import cv2
import numpy as np
from PIL import Image

DEFAULT_KEY = b"0123456789abcdef0123456789abcdef"   # watermarking_embed_test.py:14


def encrypt(text: str, key: bytes = DEFAULT_KEY, iv: bytes = b"\x07" * 16) -> bytes:
    from cryptography.hazmat.primitives import padding
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    padder = padding.PKCS7(128).padder()
    data = padder.update(text.encode("utf-8")) + padder.finalize()
    enc = Cipher(algorithms.AES(key), modes.CBC(iv)).encryptor()
    return iv + enc.update(data) + enc.finalize()          # IV || ciphertext, as encryption.py:8-40


def decrypt(blob: bytes, key: bytes = DEFAULT_KEY) -> str:
    from cryptography.hazmat.primitives import padding
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    dec = Cipher(algorithms.AES(key), modes.CBC(blob[:16])).decryptor()
    data = dec.update(blob[16:]) + dec.finalize()
    unp = padding.PKCS7(128).unpadder()
    return (unp.update(data) + unp.finalize()).decode("utf-8")


def qr_png(payload: bytes, size: int = 1000) -> bytes:
    """text_to_qrcode: base64 of the bytes -> QR (ECC H) -> size x size image -> PNG bytes
    (embed_watermark_page.py:471-490)."""
    p = cv2.QRCodeEncoder_Params()
    p.correction_level = cv2.QRCODE_ENCODER_CORRECT_LEVEL_H
    qr = cv2.QRCodeEncoder_create(p).encode(base64.b64encode(payload).decode("ascii"))
    qr = cv2.resize(qr, (size, size), interpolation=cv2.INTER_NEAREST)
    buf = io.BytesIO()
    Image.fromarray(qr).save(buf, format="PNG")
    return buf.getvalue()


def decode_map(level_map: np.ndarray, scales=(4, 6, 3, 8)):
    """qrcode_to_text on an extracted (H//8, W//8) map: threshold, nearest upscale (SURVEY.md 8(f)
    rank 4; OpenCV's detector is sensitive to the module size in pixels, so a few integer factors
    are tried in turn), detect + decode, base64 -> bytes.  None if undecodable."""
    binary = np.where(level_map >= 128, 255, 0).astype(np.uint8)
    for sc in scales:
        img = np.pad(np.kron(binary, np.ones((sc, sc), np.uint8)), 16, constant_values=255)
        txt, _, _ = cv2.QRCodeDetector().detectAndDecode(img)
        if txt:
            try:
                return base64.b64decode(txt)
            except Exception:
                return None
    return None


def decode_prepared(img):
    """qrcode_to_text on an image that is already decodable (``prepare_for_decoding``'s output)."""
    txt, _, _ = cv2.QRCodeDetector().detectAndDecode(np.asarray(img.convert("L")))
    if not txt:
        return None
    try:
        return base64.b64decode(txt)
    except Exception:
        return None
