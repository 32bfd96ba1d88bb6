#!/usr/bin/env python
"""How much of a hash depends on the details nobody can check here?  ("a sensitivity table")

The oracle restates fast_image_resize's Lanczos3 u8 convolution and rustdct's DCT-II(16) from their published
algorithms; their sources are not in the reference repository and no reference test pins a hash value.  This script
hashes the same synthetic stacks under each plausible ALTERNATIVE reading and counts the hash bits that change relative
to the oracle's reading -- the size of the risk while tests/golden/reference_vectors.json is missing.

    python scripts/resize_sensitivity.py [n_stacks] [w] [h]    -> one JSON object (table rows) on stdout
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vdf_oracle as o  # noqa: E402
from tests.golden.make_reference_inputs import smooth_stack  # noqa: E402


def lanczos3(x):
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(x == 0, 1.0, np.sin(np.pi * x) / (np.pi * x) * np.sin(np.pi * x / 3) / (np.pi * x / 3))
    return np.where((x >= -3) & (x < 3), s, 0.0)


def weights(in_size, out_size=16, dtype=np.float64):
    """dense [out, in] normalised weights, as the oracle's vdfo_resize_coeffs builds them before quantising"""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    radius = 3.0 * fs
    W = np.zeros((out_size, in_size), dtype=np.float64)
    for oo in range(out_size):
        c = (oo + 0.5) * scale
        lo, hi = int(max(np.floor(c - radius), 0)), int(min(np.ceil(c + radius), in_size))
        xs = np.arange(lo, hi)
        w = lanczos3(((xs - (c - 0.5)) / fs).astype(dtype).astype(np.float64))
        if dtype == np.float32:
            w = w.astype(np.float32).astype(np.float64)
        W[oo, lo:hi] = w / w.sum()
    return W


def quantise(W, dprec=0):
    mx = W.max()
    p = 0
    for cur in range(22):
        p = cur
        if np.round(mx * (1 << (cur + 1))) >= (1 << 15):
            break
    p = max(1, p + dprec)
    return np.clip(np.round(W * (1 << p)), -32768, 32767).astype(np.int64), p


def conv(frames, K, p, axis):
    """frames [..., H, W] int64; K [16, in]; integer convolution along axis (-1 horizontal, -2 vertical), rounded to u8"""
    if axis == -1:
        acc = frames @ K.T
    else:
        acc = np.einsum("oy,...yx->...ox", K, frames)
    return np.clip((acc + (1 << (p - 1))) >> p, 0, 255)


def resize(frames, variant):
    f = frames.astype(np.int64)
    h, w = f.shape[-2:]
    if variant == "pillow":
        from PIL import Image

        return np.stack([np.asarray(Image.fromarray(fr).resize((16, 16), Image.LANCZOS)) for fr in frames])
    dt = np.float32 if variant == "f32_weights" else np.float64
    dp = {"precision_minus_1": -1, "precision_plus_1": 1}.get(variant, 0)
    Kh, ph = quantise(weights(w, dtype=dt), dp)
    Kv, pv = quantise(weights(h, dtype=dt), dp)
    if variant == "vertical_first":
        return conv(conv(f, Kv, pv, -2), Kh, ph, -1).astype(np.uint8)
    if variant == "no_u8_between_passes":
        acc = np.einsum("oy,...yx->...ox", Kv, f @ Kh.T)
        return np.clip((acc + (1 << (ph + pv - 1))) >> (ph + pv), 0, 255).astype(np.uint8)
    return conv(conv(f, Kh, ph, -1), Kv, pv, -2).astype(np.uint8)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 180
    variants = ["numpy_restatement", "vertical_first", "precision_minus_1", "precision_plus_1", "f32_weights", "no_u8_between_passes", "pillow"]
    diff = {v: 0 for v in variants}
    pix = {v: 0 for v in variants}
    stacks_hit = {v: 0 for v in variants}
    dct_diff = near = 0
    from scipy.fft import dctn

    for s in range(n):
        st = smooth_stack(w, h, 50_000 + s)
        base_small = np.stack([o.resize_lanczos3(fr) for fr in st])
        base, coefs = o.hash_from_small(base_small, want_coefs=True)
        for v in variants:
            sm = resize(st, v)
            hv = o.hash_from_small(sm)
            d = int(np.unpackbits((hv ^ base).view(np.uint8)).sum())
            diff[v] += d
            pix[v] += int((sm != base_small).sum())
            stacks_hit[v] += d > 0
        # the DCT: scipy's (a different operation order) on the same cube; the cube is [t][row][col], the hash uses [t][x=col][y=row]
        c2 = dctn(base_small.astype(np.float64).transpose(0, 2, 1) - 128.0, type=2, norm=None) / 8.0
        a, b = coefs[:10, :10, :10] > 0, c2[:10, :10, :10] > 0
        dct_diff += int((a != b).sum())
        near += int((np.abs(coefs[:10, :10, :10]) < 1e-6).sum())
    out = {"stacks": n, "size": [w, h], "bits_per_stack": 1000,
           "resize_variants": {v: {"hash_bits_changed": diff[v], "rate": diff[v] / (n * 1000.0), "stacks_with_a_change": stacks_hit[v],
                                   "resized_pixels_changed": pix[v], "pixel_rate": pix[v] / (n * 4096.0)} for v in variants},
           "dct_scipy_vs_split_radix": {"signs_changed": dct_diff, "rate": dct_diff / (n * 1000.0), "coefficients_below_1e-6": near}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
