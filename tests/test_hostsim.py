"""The kernels' per-block arithmetic (thatsmyface_b200/csrc/tmf_math.cuh and
tmf_fast.cuh compiled for the host by tests/hostsim) against the oracle.  This
is how the CUDA math is checked in the build container, which has no GPU; the
`-m gpu` suite repeats the same checks through the real library."""
import numpy as np
import pytest

import hostsim_util as H
from conftest import golden_names
from oracle import wm_oracle as O
from oracle.make_golden import natural_like, regions

ARRAY_CASES = [n for n in golden_names() if not n.startswith(("pil_", "bs")) and n != "subblock_5x7"]
BS_CASES = [n for n in golden_names() if n.startswith("bs")]


def _tie(S, rel=1e-4):
    s0, s1 = S[..., 0], S[..., 1]
    return ((s0 - s1) <= rel * np.maximum(s0, 1e-30)) & (s0 > 0)


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("name", ARRAY_CASES)
def test_embed_extract_math_vs_reference_vectors(golden, name, mode):
    g = golden(name)
    out, sig, _ = H.embed(g["rgb"], g["wm"], mode=mode)
    d = np.abs(out.astype(int) - g["ref_out"].astype(int))
    bad = np.repeat(np.repeat(_tie(g["ref_S"]), 8, 0), 8, 1)
    d[: bad.shape[0], : bad.shape[1]][bad] = 0
    assert d.max() <= 1
    s0 = g["ref_S"][..., 0]
    assert (np.abs(sig - s0) <= 1e-5 * s0 + 1e-12).all()
    ext = H.extract(g["ref_out"], g["rgb"], mode=mode)
    de = np.abs(ext.astype(int) - g["ref_ext"].astype(int))
    assert de.max() <= 1
    decided = np.abs(g["ref_ext"].astype(int) - 128) > 1
    assert np.array_equal((ext >= 128)[decided], (g["ref_ext"] >= 128)[decided])


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("kind", ["random", "natural", "regions", "gray"])
def test_embed_extract_math_vs_oracle(kind, mode):
    rng = np.random.default_rng(5)
    img = {"random": lambda: rng.integers(0, 256, (128, 160, 3), dtype=np.uint8),
           "natural": lambda: natural_like(136, 200, 2),
           "regions": lambda: regions(128, 128, 4),
           "gray": lambda: np.repeat(natural_like(64, 64, 9)[:, :, 1:2], 3, 2)}[kind]()
    h, w = img.shape[:2]
    wm = (rng.integers(0, 2, (h // 8, w // 8)) * 255).astype(np.uint8)
    wm[0] = rng.integers(0, 256, w // 8)
    taps = {}
    ref = O.embed_array(img, wm, taps=taps)
    out, sig, iters = H.embed(img, wm, mode=mode)
    d = np.abs(out.astype(int) - ref.astype(int))
    bad = np.repeat(np.repeat(_tie(taps["S"]), 8, 0), 8, 1)
    d[bad] = 0
    assert d.max() <= 1
    if kind in ("random", "natural"):
        assert (d > 0).mean() <= 2e-3
    assert (iters % 100).max() <= (18 if mode == 1 else 12)
    rext = O.extract_array(ref, img)
    ext = H.extract(ref, img, mode=mode)
    assert np.abs(ext.astype(int) - rext.astype(int)).max() <= 1
    # own round trip recovers the bits wherever the block does not clip at white
    ext2 = H.extract(out, img, mode=mode)
    ok = O.to_blocks(O.rgb_to_ycbcr(img)[:, :, 0]).max(axis=(2, 3)) <= 0.9
    ok[0] = False
    assert np.array_equal((ext2 >= 128)[ok], (wm >= 128)[ok])


@pytest.mark.parametrize("alpha", [0.5, 1.0])
def test_alpha_range(alpha):
    img = natural_like(64, 96, 12)
    wm = np.random.default_rng(1).integers(0, 256, (8, 12), dtype=np.uint8)
    ref = O.embed_array(img, wm, alpha)
    for mode in (0, 1):
        out, _, _ = H.embed(img, wm, alpha, mode)
        assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1
        assert np.abs(H.extract(ref, img, alpha, mode).astype(int) - O.extract_array(ref, img, alpha).astype(int)).max() <= 1


def test_jacobi_svd_math():
    rng = np.random.default_rng(7)
    D = O.dct_blocks(O.to_blocks(O.rgb_to_ycbcr(regions(128, 128, 2))[:, :, 0])).reshape(-1, 8, 8)
    D = np.concatenate([D, (rng.normal(size=(256, 8, 8)) * 10.0 ** rng.integers(-12, 12, (256, 1, 1))).astype(np.float32)])
    AV, V, S, sweeps = H.svd(D)
    Sref = np.linalg.svd(D.astype(np.float64), compute_uv=False)
    s0 = np.maximum(Sref[:, :1], 1e-300)
    assert (np.abs(-np.sort(-S, axis=1) - Sref) <= 1e-5 * s0).all()
    assert np.abs(np.einsum("nki,nkj->nij", V, V) - np.eye(8)).max() <= 5e-6
    assert (np.abs(np.einsum("nik,njk->nij", AV, V) - D).reshape(len(D), -1).max(1) <= 2e-6 * s0[:, 0]).all()
    assert sweeps.max() <= 10


def test_dct_and_colour_math():
    rng = np.random.default_rng(8)
    b = rng.random((500, 8, 8), dtype=np.float32)
    assert np.abs(H.dct(b) - O.dct_blocks(b)).max() <= 2e-6
    assert np.abs(H.dct(O.dct_blocks(b), inverse=True) - b).max() <= 2e-6
    rgb = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    assert np.array_equal(H.rgb2ycc(rgb), O.rgb_to_ycbcr(rgb))


_QR_CASE = {}


def _qr_case():
    """1080p natural-like image + a 40-character text as an AES'd QR watermark, through the
    oracle once (about 6 s)."""
    if not _QR_CASE:
        import qr_util as Q

        text = "Test" * 10
        png = Q.qr_png(Q.encrypt(text))
        rgb = natural_like(1080, 1920, 21)
        wm = np.array(O.resize_watermark(png, 135, 240, True))
        ref = O.embed_array(rgb, wm)
        _QR_CASE.update(text=text, png=png, rgb=rgb, wm=wm, ref=ref, ref_ext=O.extract_array(ref, rgb))
    return _QR_CASE


@pytest.mark.parametrize("mode", [0, 1])
def test_qr_payload_survives_embed_extract(mode):
    """north_star: QR payload bit-exact.  text -> AES -> base64 -> QR -> PNG ->
    resize_watermark(preserve_ratio) -> embed -> extract -> QR decode -> AES."""
    import qr_util as Q

    c = _qr_case()
    ref_payload = Q.decode_map(c["ref_ext"])
    assert ref_payload is not None and Q.decrypt(ref_payload) == c["text"]      # the reference path itself works
    out, _, _ = H.embed(c["rgb"], c["wm"], mode=mode)
    assert np.abs(out.astype(int) - c["ref"].astype(int)).max() <= 1
    got = Q.decode_map(H.extract(out, c["rgb"], mode=mode))
    assert got == ref_payload and Q.decrypt(got) == c["text"]
    # cross paths decode to the same bytes
    assert Q.decode_map(O.extract_array(out, c["rgb"])) == ref_payload
    assert Q.decode_map(H.extract(c["ref"], c["rgb"], mode=mode)) == ref_payload


@pytest.mark.parametrize("bs", [4, 6, 10, 12, 14, 16])
def test_other_block_sizes_math_vs_oracle(bs):
    """the UI's other block sizes (embed_watermark_page.py:324-331)."""
    rng = np.random.default_rng(bs)
    img = natural_like(7 * bs + 1, 9 * bs + 3, bs)            # ragged: strips on both sides
    wm = np.where(rng.random((7, 9)) < 0.4, 0, rng.integers(1, 256, (7, 9))).astype(np.uint8)
    for alpha in (0.1, 0.6):
        taps = {}
        ref = O.embed_array(img, wm, alpha, bs, taps=taps)
        out = H.embed_n(img, wm, alpha, bs)
        d = np.abs(out.astype(int) - ref.astype(int))
        assert d.max() <= 1, (bs, alpha, d.max())
        rext = O.extract_array(ref, img, alpha, bs)
        ext = H.extract_n(ref, img, alpha, bs)
        assert np.abs(ext.astype(int) - rext.astype(int)).max() <= 1
        decided = np.abs(rext.astype(int) - 128) > 1
        assert np.array_equal((ext >= 128)[decided], (rext >= 128)[decided])
    flat = np.zeros((2 * bs, 2 * bs, 3), np.uint8)              # sigma0 = 0: mark lands on DC
    wmf = np.full((2, 2), 255, np.uint8)
    assert np.abs(H.embed_n(flat, wmf, 0.1, bs).astype(int) - O.embed_array(flat, wmf, 0.1, bs).astype(int)).max() <= 1


@pytest.mark.parametrize("name", BS_CASES)
def test_other_block_sizes_math_vs_reference_vectors(golden, name):
    g = golden(name)
    bs, alpha = int(g["bs"]), float(g["alpha"])
    out = H.embed_n(g["rgb"], g["wm"], alpha, bs)
    assert np.abs(out.astype(int) - g["ref_out"].astype(int)).max() <= 1
    ext = H.extract_n(g["ref_out"], g["rgb"], alpha, bs)
    assert np.abs(ext.astype(int) - g["ref_ext"].astype(int)).max() <= 1


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("family", ["two_pixels", "two_rectangles", "checker", "flat_noise"])
def test_adversarial_block_families(family, mode):
    """Near-tied / rank-deficient / tiny-gap blocks.  Every block whose two largest singular
    values differ by >= 1e-4 relative must match the oracle to 1 LSB (both alphas); only
    exact ties, where the reference's own u0 v0^T is LAPACK's arbitrary pick inside a 2-D
    singular subspace, are excluded.  sigma0 (extract) is well defined even there."""
    rng = np.random.default_rng(77)
    n = 2500
    B = np.zeros((n, 8, 8), np.uint8)
    for k in range(n):
        if family == "two_pixels":
            a = int(rng.integers(1, 256)); b = min(255, max(1, a + int(rng.integers(-3, 4))))
            i, j = rng.choice(8, 2, replace=False); p, q = rng.choice(8, 2, replace=False)
            B[k, i, p] = a; B[k, j, q] = b
        elif family == "two_rectangles":
            a, b = rng.integers(1, 256, 2); r, c = rng.integers(1, 8, 2)
            B[k, :r, :c] = a; B[k, r:, c:] = b
        elif family == "checker":
            base = (np.add.outer(np.arange(8), np.arange(8)) % 2) * int(rng.integers(1, 256))
            B[k] = np.clip(base + rng.integers(0, 3, (8, 8)), 0, 255)
        else:
            B[k] = np.clip(int(rng.integers(0, 256)) + rng.integers(-2, 3, (8, 8)), 0, 255)
    img = np.repeat(B.transpose(1, 0, 2).reshape(8, 8 * n)[:, :, None], 3, axis=2).copy()
    wm = rng.integers(1, 256, (1, n), dtype=np.uint8)
    for alpha in (0.1, 1.0):
        taps = {}
        ref = O.embed_array(img, wm, alpha, taps=taps)
        S = taps["S"][0]
        ok = (S[:, 0] - S[:, 1]) >= 1e-4 * np.maximum(S[:, 0], 1e-30)
        out, _, _ = H.embed(img, wm, alpha, mode)
        d = np.abs(out.astype(int) - ref.astype(int)).reshape(8, n, 8, 3).max(axis=(0, 2, 3))
        assert d[ok].max() <= 1, (family, alpha, int(d[ok].max()))
        assert np.abs(H.extract(ref, img, alpha, mode).astype(int) - O.extract_array(ref, img, alpha).astype(int)).max() <= 1


def test_dominant_column_jacobi_against_float64():
    """tmf::top_column8 - the SVD step of faithful embed / extract.  Blocks with a dominant column take the
    7-rotation sweeps (certified: |p|^2 > 4 x the rest), the others the full cyclic routine; either way sigma0
    and u0 must be the float64 ones to fp32 round-off.  Covers both paths, the limit of the certificate, a
    dominant column that is not column 0, rank-1 and zero blocks."""
    from scipy.fft import dctn
    rng = np.random.default_rng(11)
    blocks = []
    for k in range(300):                                   # DCT of noise around a mean: eligible
        blocks.append(dctn((rng.uniform(0.2, 0.9) + rng.normal(0, rng.uniform(0.0, 0.08), (8, 8))).clip(0, 1), norm="ortho"))
    for k in range(300):                                   # constructed: ratio sigma1^2/sigma0^2 from 0.02 up to 0.6
        U, _ = np.linalg.qr(rng.normal(size=(8, 8)))
        G = rng.normal(size=(8, 8)) * rng.uniform(0.0, 0.25)
        V, _ = np.linalg.qr(np.eye(8) + G - G.T)
        r = rng.uniform(0.02, 0.6)
        s = np.concatenate([[1.0, np.sqrt(r * 0.9)], np.sqrt(r * 0.1 / 6) * rng.random(6)])
        D = (U * s) @ V.T * rng.uniform(0.3, 5.0)
        blocks.append(D[:, rng.permutation(8)] if k % 2 else D)   # dominant column anywhere
    for k in range(100):                                   # two comparable patches, sparse pixels: not eligible
        b = np.zeros((8, 8))
        r_, c_ = rng.integers(1, 8, 2)
        b[:r_, :c_] = rng.uniform(0.1, 1); b[r_:, c_:] = rng.uniform(0.1, 1)
        blocks.append(dctn(b, norm="ortho"))
    blocks.append(np.outer(rng.normal(size=8), rng.normal(size=8)))   # rank 1
    blocks.append(np.zeros((8, 8)))
    B = np.array(blocks, np.float32)
    s0, u0, sw = H.top_column(B)
    assert (sw < 100).sum() >= 350 and (sw >= 100).sum() >= 100, "both paths must be exercised"
    assert sw[sw < 100].max() <= 5, f"dominant-column path took {sw[sw < 100].max()} sweeps"
    for k in range(len(B) - 1):
        U, S, _ = np.linalg.svd(B[k].astype(np.float64))
        assert abs(s0[k] - S[0]) <= 1e-6 * S[0], (k, s0[k], S[0], sw[k])
        if S[1] < 0.9 * S[0]:                              # u0 is defined up to sign
            assert min(np.abs(u0[k] - U[:, 0]).max(), np.abs(u0[k] + U[:, 0]).max()) <= 2e-4 / (1 - S[1] / S[0]), (k, sw[k])
    assert s0[-1] == 0.0
