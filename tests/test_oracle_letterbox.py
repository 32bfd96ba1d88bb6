"""Pins the ORACLE's letterbox crop-detect against the reference's 17 exact Crop KATs
(vid_dup_finder_common/src/video_frames_gray.rs:225-458).  Expected values are the reference's
`Crop::from_edge_offsets((w,h), left, right, top, bottom)` arguments, verbatim."""
import numpy as np
import pytest

from oracle import vdf_oracle as o

BW, ANY = o.LB_BLACKWHITE, o.LB_ANYCOLOUR


def img(w, h, pix):
    return np.array(pix, dtype=np.uint8).reshape(h, w)


WHITE = img(3, 3, [255] * 9)
BLACK = img(3, 3, [0] * 9)
GRAY = img(3, 3, [127, 127, 127, 127, 0, 127, 127, 127, 127])
THRESH = img(3, 3, [120, 130, 120, 130, 0, 130, 120, 130, 120])
ONEPIX = img(3, 3, [0, 0, 0, 0, 127, 0, 0, 0, 0])
TOPCORNER = img(3, 3, [127, 0, 0, 0, 0, 0, 0, 0, 0])
RIGHTEDGE = img(3, 3, [0, 0, 200, 0, 0, 120, 0, 0, 100])
BR2 = img(3, 3, [0, 0, 0, 0, 127, 0, 0, 0, 127])
BOTTOM2 = img(5, 6, [0, 0, 0, 0, 0,
                     0, 255, 255, 255, 0,
                     0, 255, 255, 255, 0,
                     0, 255, 255, 255, 0,
                     0, 0, 0, 0, 0,
                     0, 0, 0, 0, 0])

KATS = [
    # (name, image, mode, tol, expected (l, r, t, b))                       reference test
    ("white_bw", WHITE, BW, 1, (0, 0, 0, 0)),       # test_letterbox_crop_white_img_finds_no_crop :225
    ("white_any", WHITE, ANY, 1, (0, 0, 0, 0)),
    ("black_bw", BLACK, BW, 1, (0, 0, 0, 0)),       # test_letterbox_crop_black_img_finds_no_crop :254
    ("black_any", BLACK, ANY, 1, (0, 0, 0, 0)),
    ("gray_bw", GRAY, BW, 1, (0, 0, 0, 0)),         # test_letterbox_crop_any_colour_gray :282
    ("gray_any", GRAY, ANY, 1, (1, 1, 1, 1)),
    ("thresh_9", THRESH, ANY, 9, (0, 0, 0, 0)),     # test_letterbox_crop_any_threshold :310
    ("thresh_10", THRESH, ANY, 10, (1, 1, 1, 1)),
    ("onepix_bw", ONEPIX, BW, 10, (1, 1, 1, 1)),    # test_letterbox_crop_onepix :337
    ("onepix_any", ONEPIX, ANY, 1, (1, 1, 1, 1)),
    ("topcorner_bw", TOPCORNER, BW, 10, (0, 2, 0, 2)),   # test_letterbox_crop_topcorner :365
    ("topcorner_any", TOPCORNER, ANY, 1, (0, 2, 0, 2)),
    ("rightedge_bw", RIGHTEDGE, BW, 10, (2, 0, 0, 0)),   # test_letterbox_crop_rightedge :393
    ("rightedge_any", RIGHTEDGE, ANY, 1, (2, 0, 0, 0)),
    ("br2_bw", BR2, BW, 10, (1, 0, 1, 0)),          # test_letterbox_crop_bottom_right_2pix :421
    ("br2_any", BR2, ANY, 1, (1, 0, 1, 0)),
    ("bottom2_any", BOTTOM2, ANY, 1, (1, 1, 1, 2)),  # test_letterbox_crop_2pix_bottom :441
]


@pytest.mark.parametrize("name,image,mode,tol,exp", KATS, ids=[k[0] for k in KATS])
def test_letterbox_kat(name, image, mode, tol, exp):
    assert o.letterbox_frame(image, mode, tol) == exp


def test_mode_tie_takes_last_maximum():
    """Iterator::max_by_key returns the LAST maximum (video_frames_gray.rs:82-87): a top row holding five
    0s and five 200s has mode 200, so a tolerance that only reaches from 200 decides the strip."""
    im = np.full((4, 10), 90, np.uint8)
    im[0, :5] = 0
    im[0, 5:] = 200
    im[0, 4] = 190  # 4x0, 1x190, 5x200 -> mode 200 ; |200-190|<=16 -> 6/10 match -> not letterbox
    assert o.letterbox_frame(im, ANY, 16)[2] == 0
    im[0, :] = 200
    im[0, 0] = 0  # 9/10 = 0.9 is NOT > 0.9 (strict)
    assert o.letterbox_frame(im, ANY, 16)[2] == 0


def test_ratio_test_equals_integer_form():
    """proportion > 0.9 in f64 (video_frames_gray.rs:97-100) == 10*count > 9*len for every strip length the
    path can see; the GPU kernel uses the integer form."""
    for length in list(range(1, 400)) + [720, 1080, 1920, 2160, 3840, 4096, 7680]:
        for count in {0, length, (9 * length) // 10, (9 * length) // 10 + 1, max((9 * length) // 10 - 1, 0)}:
            assert (count / length > 0.9) == (10 * count > 9 * length), (count, length)


def test_cropdetect_uses_frames_0_and_8_with_per_side_min():
    """cropdetect_letterbox: step_by(8).take(8) (video_frames_gray.rs:201-210), union = per-side min (crop.rs:53-68)"""
    rng = np.random.default_rng(5)
    frames = rng.integers(60, 200, (16, 40, 64), dtype=np.uint8)
    frames[:, :6, :] = 16   # 6-row top bar everywhere
    frames[:, -4:, :] = 16  # 4-row bottom bar
    frames[8, 4:6, :] = rng.integers(60, 200, (2, 64), dtype=np.uint8)  # frame 8 only has a 4-row top bar
    frames[3, :, :] = 16    # frames other than 0 and 8 are never looked at
    st, crop = o.cropdetect_letterbox(frames)
    assert st == 0 and crop == (0, 0, 4, 4)
    # a uniform frame converges from both sides -> that frame contributes no crop at all (:119-127)
    frames[8] = 77
    assert o.cropdetect_letterbox(frames)[1] == (0, 0, 0, 0)
