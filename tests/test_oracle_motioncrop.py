"""The reference's seven motion-crop tests (vid_dup_finder_common/src/motioncrop/test.rs:9-225) against the CPU restatement
of `Cropdetect::Motion` (oracle/motioncrop_oracle.py).  SURVEY.md section 8(f) N4: these tests pin the oracle
that csrc/motion.cu is compared with (tests/test_gpu_hashing.py::test_motion_crop_matches_oracle).  Frames are repeated twice, as
`util_generate_frames` does (test.rs:230-242); expected crops are (left, right, top, bottom)."""
import numpy as np
import pytest

from oracle import motioncrop_oracle as m


def _frames(w, h, *pix):
    return [np.array(p, dtype=np.uint8).reshape(h, w) for p in pix]


CASES = {
    "nocrop": (3, 3, [[255] * 9, [255] * 9], (0, 0, 0, 0)),  # test.rs:9-31
    "letterbox_static": (5, 6, [  # test.rs:35-63
        [0, 0, 0, 0, 0, 0, 255, 255, 255, 0, 0, 255, 255, 255, 0, 0, 255, 255, 255, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]] * 2, (1, 1, 1, 2)),
    "2pixsquareinthemiddle": (4, 4, [  # test.rs:65-90
        [255, 220, 220, 255, 220, 80, 80, 220, 220, 80, 80, 220, 255, 255, 255, 255],
        [255, 220, 220, 255, 220, 27, 27, 220, 220, 27, 27, 220, 255, 255, 255, 255]], (1, 1, 1, 1)),
    "prefer_bigger_region": (4, 8, [  # test.rs:92-124
        [255, 220, 220, 255, 220, 80, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 80, 80, 220, 220, 80, 80, 220, 255, 255, 255, 255],
        [255, 220, 220, 255, 220, 20, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 20, 20, 220, 220, 20, 20, 220, 255, 255, 255, 255]], (1, 1, 5, 1)),
    "prefer_upper_region": (4, 8, [  # test.rs:126-158
        [255, 220, 220, 255, 220, 80, 80, 220, 220, 255, 80, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 80, 80, 220, 220, 80, 80, 220, 255, 255, 255, 255],
        [255, 220, 220, 255, 220, 20, 255, 220, 220, 20, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 20, 20, 220, 220, 20, 20, 220, 255, 255, 255, 255]], (1, 1, 1, 5)),
    "detect_topleft": (4, 8, [  # test.rs:160-192
        [80, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 220, 255, 255],
        [20, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255]], (0, 3, 0, 7)),
    "detect_botright": (4, 8, [  # test.rs:194-225
        [255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 20, 20, 255, 255, 20, 20],
        [255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 255, 220, 255, 255, 255, 255,
         255, 220, 220, 255, 220, 255, 255, 220, 220, 255, 40, 20, 255, 255, 20, 40]], (2, 0, 6, 0)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_reference_motioncrop_cases(name):
    w, h, pix, want = CASES[name]
    assert m.cropdetect_motion(_frames(w, h, *pix)) == want


def test_fewer_than_two_frames_is_none():  # autocrop_frames.rs:46-48
    assert m.cropdetect_motion([np.zeros((4, 4), np.uint8)]) is None
    assert m.cropdetect_motion([]) is None


def test_building_blocks():
    img = np.zeros((9, 9), np.uint8)
    img[4, 4] = 255
    assert m.dilate(img, 2).sum() == 25 * 255 and m.erode(m.dilate(img, 2), 2)[4, 4] == 255  # L-inf balls are squares
    assert m.erode(np.full((5, 5), 255, np.uint8), 3).min() == 255  # the image border is not background
    lab = m.connected_components8(np.array([[1, 0, 1], [0, 1, 0], [0, 0, 0], [1, 0, 0]], np.uint8))
    assert lab.tolist() == [[1, 0, 1], [0, 1, 0], [0, 0, 0], [2, 0, 0]]  # diagonal neighbours connect; raster-order labels
    assert m.u16_to_u8(np.array([0, 128, 129, 65535], np.uint16)).tolist() == [0, 0, 1, 255]
    flat = m.blur(np.full((6, 7), 200, np.uint8), 2.0)
    assert (flat == 200).all()  # re-normalised taps: a constant image stays constant up to the border
    assert m.stretch_contrast(np.array([[10, 60, 110]], np.uint8), 10, 110).tolist() == [[0, 127, 255]]
