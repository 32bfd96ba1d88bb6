"""ORACLE hashing path: PARITY UNPINNED against the reference (no golden hash values exist in its tests and
the arithmetic lives in un-vendored crates), so it is anchored three ways:
  * the resize restatement against Pillow's LANCZOS (same Pillow-SIMD lineage as fast_image_resize),
  * the split-radix DCT against scipy.fft (sign-equivalent f64 DCT-II),
  * the reference's end-to-end expectation on its own example clips (examples/example.rs:77-82).
"""
import json
import os

import numpy as np
import pytest
import scipy.fft as sf

from oracle import vdf_oracle as o
from tests import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dct16_matches_scipy():
    rng = np.random.default_rng(0)
    for _ in range(50):
        x = rng.integers(-128, 128, 16).astype(np.float64)
        assert np.abs(o.dct2_16(x) - sf.dct(x, type=2) / 2).max() < 1e-10  # unnormalised: X_k = sum x_n cos(...)


def test_dct3d_matches_scipy_and_axis_order():
    rng = np.random.default_rng(1)
    cube = rng.integers(-128, 128, (16, 16, 16)).astype(np.float64)
    got = o.dct3d(cube)
    assert np.abs(got - sf.dctn(cube, type=2) / 8).max() < 1e-8
    # DC term = plain sum (raw_dct_ops.rs:138-139: normalisation is commented out)
    assert abs(got[0, 0, 0] - cube.sum()) < 1e-6


def test_static_stack_gives_exact_zero_temporal_coefficients():
    """SURVEY 'hard parts': a static stack must give exact 0.0 (bit 0) for every t>0 coefficient."""
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (16, 16), dtype=np.uint8)
    small = np.broadcast_to(frame, (16, 16, 16)).copy()
    h, coefs = o.hash_from_small(small, want_coefs=True)
    assert np.count_nonzero(coefs[1:]) == 0
    bits = np.unpackbits(h.view(np.uint8), bitorder="little")
    assert bits[100:].sum() == 0  # bits t*100 + x*10 + y with t >= 1
    # mirror-symmetric frames: odd-x coefficients are exactly zero too
    sym = np.concatenate([frame[:, :8], frame[:, 7::-1]], axis=1)
    _, c2 = o.hash_from_small(np.broadcast_to(sym, (16, 16, 16)).copy(), want_coefs=True)
    assert np.count_nonzero(c2[0, 1::2, :]) == 0


def test_hash_bit_layout():
    """bit b = t*100 + x*10 + y <- coef[t][x][y] > 0 with m[t][x=col][y=row] (dct_3d.rs:40-44,55-66);
    Lsb0 packing, bits 1000..1023 zero (video_hash.rs:63-70)."""
    rng = np.random.default_rng(3)
    small = rng.integers(0, 256, (16, 16, 16), dtype=np.uint8)
    h, coefs = o.hash_from_small(small, want_coefs=True)
    m = small.astype(np.float64).transpose(0, 2, 1) - 128.0  # [t][col][row]
    ref = sf.dctn(m, type=2) / 8
    assert np.abs(coefs - ref).max() < 1e-8
    want = (ref[:10, :10, :10] > 0).reshape(-1)
    bits = np.unpackbits(h.view(np.uint8), bitorder="little")
    safe = np.abs(ref[:10, :10, :10]).reshape(-1) > 1e-6
    assert np.array_equal(bits[:1000][safe], want[safe].astype(np.uint8))
    assert bits[1000:].sum() == 0


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (854, 480), (640, 360), (256, 144), (100, 37), (16, 16), (7, 5)])
def test_resize_matches_pillow_lanczos(w, h):
    """the i16-coefficient restatement vs PIL.Image.resize(LANCZOS): <=1 LSB, rare."""
    from PIL import Image

    ndiff = tot = 0
    for seed in range(3):
        img = synth.smooth_frame(w, h, seed)
        got = o.resize_lanczos3(img)
        ref = np.asarray(Image.fromarray(img).resize((16, 16), Image.LANCZOS))
        d = np.abs(got.astype(int) - ref.astype(int))
        assert d.max() <= 1
        ndiff += int((d > 0).sum())
        tot += d.size
    assert ndiff <= max(2, tot // 50)


def test_resize_coefficient_tables():
    """fast_image_resize Normalizer16: i16 coefficients, rows sum to ~2^precision, precision 21/20 for
    1920/1080 -> 16."""
    for size, prec_want in ((1920, 21), (1080, 20)):
        bounds, k, prec = o.resize_coeffs(size)
        assert prec == prec_want
        assert np.abs(k.sum(1) - (1 << prec)).max() < 64
        assert k.max() < 2**15 and bounds[:, 0].min() == 0 and (bounds[:, 0] + bounds[:, 1]).max() == size
        assert bounds[:, 1].max() <= k.shape[1] == 2 * int(np.ceil(3 * size / 16)) + 1
    _, k, prec = o.resize_coeffs(16)  # same size: identity kernel
    assert prec == 14 and all(k[i].max() == 1 << 14 and np.count_nonzero(k[i]) == 1 for i in range(16))


def test_crop_window_is_a_standalone_image():
    """The reference materialises the cropped frame, then resizes with a zero crop box
    (video_hash_builder.rs:198-201, video_hash.rs:57-59)."""
    img = synth.smooth_frame(320, 200, 7)
    l, t, cw, ch = 13, 21, 250, 150
    a = o.resize_lanczos3(img, crop=(l, t, cw, ch))
    b = o.resize_lanczos3(np.ascontiguousarray(img[t:t + ch, l:l + cw]))
    assert np.array_equal(a, b)


def test_status_codes():
    """NotEnoughFrames for < 16 frames (dct_3d.rs:47-52, video_hash.rs:61); VidProc on a size mismatch
    (video_hash_builder.rs:169-186)."""
    fr = np.stack([synth.smooth_frame(64, 48, s) for s in range(16)])
    assert o.hash_stack(fr, 1)[0] == o.OK
    assert o.hash_stack(fr[:15], 1)[0] == o.NOT_ENOUGH_FRAMES
    assert o.hash_stack(fr[:0].reshape(0, 48, 64), 1)[0] == o.NOT_ENOUGH_FRAMES
    dims = [(64, 48)] * 16
    dims[9] = (64, 50)
    assert o.hash_stack(fr, 1, frame_dims=dims)[0] == o.VIDPROC
    # more than 16 frames: only the first 16 are used (video_hash_builder.rs:164 take(DCT_SIZE))
    more = np.concatenate([fr, fr[:3]])
    assert np.array_equal(o.hash_stack(more, 1)[1], o.hash_stack(fr, 1)[1])


def test_reference_example_clips_group_as_the_reference_expects():
    """examples/example.rs:77-82 / lib.rs:26-62 on the decodable clips: cats together, dogs together, no
    cat-dog match at DEFAULT_SEARCH_TOLERANCE; a letterboxed, rescaled copy still matches its source."""
    z = np.load(os.path.join(GOLD, "ref_clips_gray.npz"))
    gold = json.load(open(os.path.join(GOLD, "ref_clips_hashes.json")))
    keys = [n.replace(".", "_") for n in z["names"]] + ["cat_1_letterboxed"]
    H = []
    for k in keys:
        st, h, crop, _ = o.hash_stack(z[k], 1)
        assert st == o.OK
        assert [format(int(w), "016x") for w in h] == gold[k]["hash_words_hex"]
        assert list(crop) == gold[k]["crop_lrtb"]
        H.append(h)
    assert gold["cat_1_letterboxed"]["crop_lrtb"] == [32, 32, 18, 18]
    H = np.stack(H)
    dur = np.array(list(z["durations"]) + [45], np.uint32)
    order = o.sort_order(dur, keys)
    gp, mm = o.search_self(H[order], dur[order], o.tolerance_int(0.35))
    groups = sorted(sorted(keys[order[i]] for i in mm[gp[g]:gp[g + 1]]) for g in range(len(gp) - 1))
    assert groups == [["cat_1_letterboxed", "cat_1_mp4", "cat_3_webm"], ["dog_1_mp4", "dog_3_webm"]]
