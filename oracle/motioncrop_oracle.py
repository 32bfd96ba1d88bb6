"""CPU restatement of `Cropdetect::Motion` (TEST INFRASTRUCTURE for SURVEY.md section 8(f) N4: the product path is
csrc/motion.cu, compared with this file bit for bit by tests/test_gpu_hashing.py::test_motion_crop_matches_oracle).

Follows vid_dup_finder_common/src/motioncrop/{autocrop_frames.rs:36-316, darkest_frame.rs:19-111, frame_change.rs:15-132,
utils.rs:8-131} and crop.rs:32-199.  The arithmetic of four un-vendored crates is restated from their published
behaviour (no source under /root/reference, `image = "0.25"`, `imageproc = "0.25"`, no Cargo.lock):
  * imageproc::contrast::stretch_contrast_mut - u16 integer map of [lo, hi] onto [0, 255], truncating;
  * imageproc::contrast::threshold_mut(.., Binary) - p > t ? 255 : 0;
  * imageproc::morphology::{open, close}(LInf, k) - through the L-inf distance transform: dilate = within Chebyshev
    distance k of a foreground pixel, erode = within distance k of a BACKGROUND pixel (the image border is not background);
  * imageproc::region_labelling::connected_components(Eight, background 0) - labels 1.. in raster order of first appearance;
  * image::imageops::blur(sigma) - separable Gaussian resampling filter, support 2 sigma (taps floor(c - 2 sigma) ..
    ceil(c + 2 sigma), clamped to the image and re-normalised), f32, vertical pass into f32, horizontal pass rounded to
    nearest and clamped;
  * image's u16 -> u8 sample conversion ((c + 128) / 257).
PINNED on the reference's own seven tests (motioncrop/test.rs:9-225, tests/test_oracle_motioncrop.py) as far as they reach:
their frames are at most 5 x 8 pixels, and all seven still pass with the blur or the close replaced by the identity - they
pin the letterbox / region / selection logic, not those two steps.  PARITY UNPINNED for the blur and the morphology.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import vdf_oracle as o

Crop = Tuple[int, int, int, int]  # (left, right, top, bottom), with the frame size known from context


# ---- third-party pieces --------------------------------------------------------------------------------------------
def stretch_contrast(img: np.ndarray, lo: int, hi: int) -> np.ndarray:
    p = np.clip(img.astype(np.uint16), lo, hi)
    return ((p - lo) * 255 // (hi - lo)).astype(np.uint8)


def threshold_binary(img: np.ndarray, t: int) -> np.ndarray:
    return np.where(img > t, 255, 0).astype(np.uint8)


def _chebyshev_to(mask: np.ndarray) -> np.ndarray:
    """L-inf distance of every pixel to the nearest True pixel of `mask`, saturating at 255 (two-pass chamfer with unit
    steps to the 8 neighbours, which is exact for the L-inf norm)"""
    h, w = mask.shape
    d = np.where(mask, 0, 255).astype(np.int32)
    for y in range(h):
        for x in range(w):
            v = d[y, x]
            for dy, dx in ((-1, -1), (-1, 0), (-1, 1), (0, -1)):
                yy, xx = y + dy, x + dx
                if 0 <= yy < h and 0 <= xx < w:
                    v = min(v, d[yy, xx] + 1)
            d[y, x] = v
    for y in range(h - 1, -1, -1):
        for x in range(w - 1, -1, -1):
            v = d[y, x]
            for dy, dx in ((1, 1), (1, 0), (1, -1), (0, 1)):
                yy, xx = y + dy, x + dx
                if 0 <= yy < h and 0 <= xx < w:
                    v = min(v, d[yy, xx] + 1)
            d[y, x] = v
    return np.minimum(d, 255)


def dilate(img: np.ndarray, k: int) -> np.ndarray:
    return np.where(_chebyshev_to(img != 0) <= k, 255, 0).astype(np.uint8)


def erode(img: np.ndarray, k: int) -> np.ndarray:
    return np.where(_chebyshev_to(img == 0) <= k, 0, 255).astype(np.uint8)


def morph_close(img: np.ndarray, k: int) -> np.ndarray:
    return erode(dilate(img, k), k)


def morph_open(img: np.ndarray, k: int) -> np.ndarray:
    return dilate(erode(img, k), k)


def connected_components8(img: np.ndarray) -> np.ndarray:
    """labels (u32) of the 8-connected non-zero regions, 1.. in raster order of each region's first pixel"""
    h, w = img.shape
    lab = np.zeros((h, w), np.uint32)
    nxt = 1
    for y in range(h):
        for x in range(w):
            if img[y, x] == 0 or lab[y, x]:
                continue
            stack = [(y, x)]
            lab[y, x] = nxt
            while stack:
                cy, cx = stack.pop()
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        yy, xx = cy + dy, cx + dx
                        if 0 <= yy < h and 0 <= xx < w and img[yy, xx] != 0 and not lab[yy, xx]:
                            lab[yy, xx] = nxt
                            stack.append((yy, xx))
            nxt += 1
    return lab


def _gauss_weights(n: int, sigma: float):
    """per output index: (first tap, f32 weights) of image::imageops::sample with a Gaussian of support 2 sigma, ratio 1"""
    f = np.float32
    support = f(2.0) * f(sigma)
    out = []
    for c in range(n):
        centre = (f(c) + f(0.5)) * f(1.0)
        left = int(np.clip(int(np.floor(centre - support)), 0, n - 1))
        right = int(np.clip(int(np.ceil(centre + support)), left + 1, n))
        centre = centre - f(0.5)
        ws = []
        total = f(0.0)
        for i in range(left, right):
            x = (f(i) - centre) / f(1.0)
            wv = f(1.0) / (np.sqrt(f(2.0) * f(np.pi)) * f(sigma)) * np.exp(-(x * x) / (f(2.0) * f(sigma) * f(sigma)), dtype=f)
            ws.append(f(wv))
            total = f(total + f(wv))
        out.append((left, [f(wv / total) for wv in ws]))
    return out


def blur(img: np.ndarray, sigma: float) -> np.ndarray:
    f = np.float32
    h, w = img.shape
    tmp = np.zeros((h, w), f)
    for y, (top, ws) in enumerate(_gauss_weights(h, sigma)):  # vertical pass first, into f32
        for x in range(w):
            t = f(0.0)
            for i, wv in enumerate(ws):
                t = f(t + f(img[top + i, x]) * wv)
            tmp[y, x] = t
    out = np.zeros((h, w), np.uint8)
    for x, (left, ws) in enumerate(_gauss_weights(w, sigma)):
        for y in range(h):
            t = f(0.0)
            for i, wv in enumerate(ws):
                t = f(t + tmp[y, left + i] * wv)
            t = min(max(float(t), 0.0), 255.0)
            out[y, x] = int(np.floor(t + 0.5))  # f32::round: half away from zero (t >= 0 here)
    return out


def u16_to_u8(img: np.ndarray) -> np.ndarray:
    return ((img.astype(np.uint32) + 128) // 257).astype(np.uint8)


# ---- the reference's own code ----------------------------------------------------------------------------------------
def _normalize_u16(s: np.ndarray) -> np.ndarray:  # frame_change.rs:109-132
    mn, mx = int(s.min()), int(s.max())
    if mn == mx:  # 65535 / 0 = inf; (pix - min) * inf = 0 * inf = NaN; NaN as u16 = 0
        return np.zeros_like(s, dtype=np.uint16)
    scale = 65535.0 / float(mx - mn)
    return np.clip((s.astype(np.float64) - mn) * scale, 0.0, 65535.0).astype(np.uint16)  # `as u16` truncates


def motion_mask(frames: Sequence[np.ndarray]) -> np.ndarray:  # FrameChange, frame_change.rs:15-88
    acc = np.zeros(frames[0].shape, np.uint16)
    for a, b in zip(frames[:-1], frames[1:]):
        d = np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.uint16)
        acc = acc + np.where(d >= 8, d, 0).astype(np.uint16)  # u16 `+=` (16 frames x 255 cannot overflow)
    img = u16_to_u8(_normalize_u16(acc))
    img = threshold_binary(blur(img, 2.0), 20)
    return morph_close(img, 5)


def dark_mask(frames: Sequence[np.ndarray]) -> np.ndarray:  # DarkestFrame::postprocess, darkest_frame.rs:19-70
    darkest = np.minimum.reduce([np.asarray(f, np.uint8) for f in frames])
    return threshold_binary(np.where(darkest >= 210, 0, 255).astype(np.uint8), 209)


def largest_dark_region_with_motion(dark: np.ndarray, motion: np.ndarray) -> Optional[np.ndarray]:  # darkest_frame.rs:84-110
    h = dark.shape[0]
    erode_thr = min(h // 10, 10)
    pp = morph_open(dark, erode_thr) if h > 100 else dark.copy()
    anded = (pp == 255) & (motion == 255)
    regions = connected_components8(pp)
    keep: List[int] = []
    for v in regions[anded].tolist():  # regions_in_mask: raster order, no duplicates (utils.rs:32-42)
        if v not in keep:
            keep.append(v)
    preserved = np.where(np.isin(regions, keep), regions, 0)
    labels = preserved[preserved != 0]
    if labels.size == 0:
        return None
    counts = np.bincount(labels)
    best = int(np.flatnonzero(counts == counts.max())[-1])  # Iterator::max_by: the LAST maximum (utils.rs:56-70)
    return np.where(preserved == best, 255, 0).astype(np.uint8)


def _from_frames_one(frames: Sequence[np.ndarray]) -> Optional[Crop]:  # autocrop_frames.rs:220-310
    h, w = frames[0].shape
    if len(frames) < 2:  # FrameChange::try_from_iter over zero pairs
        return None
    mask = largest_dark_region_with_motion(dark_mask(frames), motion_mask(frames))
    if mask is None:
        return None
    ys, xs = np.nonzero(mask == 255)
    if ys.size == 0:
        return None
    x, y, bw, bh = int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)
    ret = (x, w - bw - x, y, h - bh - y)  # Crop::from_topleft_and_dims (crop.rs:32-50)
    if ret == (0, 0, 0, 0):
        return ret
    e = _eroded(_eroded(ret, w, h), w, h)
    return e if e is not None else ret


def _eroded(c: Optional[Crop], w: int, h: int) -> Optional[Crop]:  # crop.rs:165-181
    if c is None:
        return None
    l, r, t, b = (v + 1 for v in c)
    if l + r >= w or t + b >= h:
        return None
    return (l, r, t, b)


def cropdetect_motion(frames: Sequence[np.ndarray]) -> Optional[Crop]:
    """MotiondetectCrop::from_frames (autocrop_frames.rs:36-218) -> (left, right, top, bottom) or None"""
    frames = [np.array(f, dtype=np.uint8, copy=True) for f in frames]
    if len(frames) < 2:
        return None
    mn, mx = min(int(f.min()) for f in frames), max(int(f.max()) for f in frames)
    if mx != 255 and mn != 0 and mn < mx:
        frames = [stretch_contrast(f, mn, mx) for f in frames]
    if any(f.shape != frames[0].shape for f in frames):
        return None
    h, w = frames[0].shape
    lb = None
    for f in frames:  # union of every frame's letterbox: per-side minimum (crop.rs:53-68)
        c = o.letterbox_frame(f)
        lb = c if lb is None else tuple(min(a, b) for a, b in zip(lb, c))
    l, r, t, b = lb
    for f in frames:  # whiten everything outside the letterbox crop
        keep = f[t:h - b, l:w - r].copy()
        f[:, :] = 255
        f[t:h - b, l:w - r] = keep
    crop_1 = _from_frames_one(frames)
    crop_2 = None
    if crop_1 is not None:
        l1, r1, t1, b1 = crop_1
        for f in frames:
            f[t1:h - b1, l1:w - r1] = 255  # clear_out_cropped_area (utils.rs:126-130)
        crop_2 = _from_frames_one(frames)
    crops = [c for c in (crop_1, crop_2) if c is not None]
    if not crops:
        return lb

    def dims(c):
        return w - c[0] - c[1], h - c[2] - c[3]

    largest = max(dims(c)[0] * dims(c)[1] for c in crops)
    ok = []
    for c in crops:
        cw, ch = dims(c)
        ar = cw / ch if cw > ch else ch / cw
        if ar <= 3.0 and float(cw * ch) > largest * 0.8:
            ok.append(c)
    if not ok:
        return lb
    return min(ok, key=lambda c: c[2])  # min_by_key(top): the first minimum
