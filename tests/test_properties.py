"""Size-independent properties of the path, checked on the kernel arithmetic compiled for
the host (tests/hostsim) with hypothesis-generated inputs; the `-m gpu` suite checks the
same properties at 1080p on the device."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import hostsim_util as H
from oracle import wm_oracle as O


def _image(seed, h, w, kind):
    rng = np.random.default_rng(seed)
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = (110 + 60 * np.sin(x / 13.0 + seed) * np.cos(y / 9.0))[..., None] + rng.normal(0, 6, (h, w, 3))
    return np.clip(img, 20, 200).astype(np.uint8)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10**6), nbh=st.integers(1, 6), nbw=st.integers(1, 6), mode=st.integers(0, 1),
       kind=st.integers(0, 1), alpha=st.sampled_from([0.1, 0.3, 1.0]))
def test_embed_then_extract_recovers_the_mark(seed, nbh, nbw, mode, kind, alpha):
    """extract(embed(x, w), x) gives back w up to the reference's own quantisation:
    the level is (sigma0' - sigma0)/alpha * 255 truncated; the embed's truncating pixel
    quantiser costs up to 1 LSB per pixel, i.e. up to 8/255 of sigma0 = 80 levels at alpha 0.1
    (typically ~40: the KATs extract white as 213..239), and 0 stays 0."""
    rgb = _image(seed, 8 * nbh, 8 * nbw, kind)
    rng = np.random.default_rng(seed + 1)
    wm = rng.integers(0, 256, (nbh, nbw), dtype=np.uint8)
    wm[rng.random(wm.shape) < 0.3] = 0
    out, _, _ = H.embed(rgb, wm, alpha, mode)
    ext = H.extract(out, rgb, alpha, mode).astype(int)
    ref_ext = O.extract_array(O.embed_array(rgb, wm, alpha), rgb, alpha).astype(int)
    assert np.abs(ext - ref_ext).max() <= 2               # two quantisers (embed, extract) of slack
    if kind == 1:                                         # no clipping at black / white
        assert (ext[wm == 0] <= 1).all()
        assert (ext <= wm.astype(int) + 1).all() and (ext >= wm.astype(int) - 85 / (alpha / 0.1) - 3).all()


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10**6), mode=st.integers(0, 1))
def test_zero_mark_is_the_colour_round_trip_and_extracts_to_zero(seed, mode):
    rgb = _image(seed, 24, 32, seed % 2)
    wm = np.zeros((3, 4), np.uint8)
    out, _, _ = H.embed(rgb, wm, 0.1, mode)
    d = out.astype(int) - rgb.astype(int)
    assert set(np.unique(d)) <= {-1, 0, 1}                # reference: {-1, 0}; +-1 LSB tolerance on top
    assert np.abs(out.astype(int) - O.ycbcr_to_rgb(O.rgb_to_ycbcr(rgb)).astype(int)).max() <= 1
    assert (H.extract(rgb, rgb, 0.1, mode) == 0).all()    # identical images: sigma difference is exactly 0


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10**6), mode=st.integers(0, 1))
def test_blocks_are_independent(seed, mode):
    """Changing one block of the input / one mark changes only that block of the output
    (what makes by-image and by-block sharding exact)."""
    rgb = _image(seed, 24, 24, 1)
    wm = np.random.default_rng(seed).integers(1, 256, (3, 3), dtype=np.uint8)
    base, _, _ = H.embed(rgb, wm, 0.1, mode)
    rgb2, wm2 = rgb.copy(), wm.copy()
    rgb2[8:16, 8:16] = 255 - rgb2[8:16, 8:16]
    wm2[0, 2] = 0
    other, _, _ = H.embed(rgb2, wm2, 0.1, mode)
    changed = np.zeros((24, 24), bool)
    changed[8:16, 8:16] = True
    changed[0:8, 16:24] = True
    assert np.array_equal(base[~changed], other[~changed])
