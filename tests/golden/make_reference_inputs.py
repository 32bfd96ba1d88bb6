#!/usr/bin/env python
"""Inputs of the parity-pinning kit (tests/golden/refgen): a deterministic set of frame stacks, single frames and
16^3 cubes, written as one binary file that the Rust program feeds to the REAL crates (fast_image_resize, rustdct,
vid_dup_finder_common, vid_dup_finder_lib).  numpy only, so that any machine produces the same bytes.

    python tests/golden/make_reference_inputs.py [out.bin]      (default tests/golden/refgen/inputs.bin, git-ignored)

File: b"VDFI" u32 n_items, then per item: u32 kind (0 stack, 1 single frame, 2 cube), u32 name_len, name, u32 w, u32 h,
u32 n_frames, n_frames*h*w bytes.  Little endian.  `items()` is imported by tests/test_reference_vectors.py, which checks
the sha256 of every item against the one the Rust program recorded."""
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KIND_STACK, KIND_FRAME, KIND_CUBE = 0, 1, 2


def smooth_stack(w: int, h: int, seed: int, static: bool = False) -> np.ndarray:
    """[16,h,w] u8: four spatio-temporal cosines + U[-8,8] noise (the generator family of tests/synth.py, numpy only)"""
    rng = np.random.default_rng([0xB200, seed])
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    tt = np.arange(16, dtype=np.float64)[:, None, None] / 16
    img = np.full((16, h, w), 128.0)
    for _ in range(4):
        a = rng.uniform(20, 50)
        fx, fy, ft = rng.uniform(-3, 3, 3)
        img += a * np.cos(2 * np.pi * (fx * xx / w + fy * yy / h + (0.0 if static else ft) * tt) + rng.uniform(0, 2 * np.pi))
    img = img + rng.integers(-8, 9, (1 if static else 16, h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def items():
    """-> list of (kind, name, array [n_frames,h,w] u8)"""
    out = []
    z = np.load(os.path.join(HERE, "ref_clips_gray.npz"))
    for k in ("cat_1_mp4", "cat_3_webm", "dog_1_mp4", "dog_3_webm", "cat_1_letterboxed"):
        out.append((KIND_STACK, "clip:" + k, np.ascontiguousarray(z[k])))
    sizes = [(64, 48), (160, 90), (161, 91), (256, 144), (320, 180), (321, 203), (427, 240), (640, 360)]
    for s in range(32):  # synthetic stacks, with bars on some so that crop detection and odd crop sizes are exercised
        w, h = sizes[s % len(sizes)]
        st = smooth_stack(w, h, s, static=(s % 11 == 10))
        rng = np.random.default_rng([7, s])
        if s % 3 == 1:
            t, b = int(rng.integers(1, h // 6)), int(rng.integers(1, h // 6))
            st[:, :t, :] = 16 + rng.integers(-4, 5, (16, t, w))
            st[:, h - b:, :] = 16 + rng.integers(-4, 5, (16, b, w))
        if s % 3 == 2:
            l, r = int(rng.integers(1, w // 6)), int(rng.integers(1, w // 6))
            st[:, :, :l] = 235
            st[:, :, w - r:] = 235
        out.append((KIND_STACK, "synthetic:%02d:%dx%d" % (s, w, h), st))
    for s, (w, h) in enumerate([(16, 16), (17, 16), (31, 47), (100, 100), (333, 77), (640, 360), (1280, 720), (1920, 1080)]):
        out.append((KIND_FRAME, "frame:%dx%d" % (w, h), smooth_stack(w, h, 1000 + s)[:1]))
    rng = np.random.default_rng(11)
    cubes = rng.integers(0, 256, (24, 16, 16, 16), dtype=np.uint8)
    cubes[0] = 128                                   # every coefficient exactly 0.0
    cubes[1] = cubes[1, 0]                           # static: every t > 0 coefficient exactly 0.0 in a butterfly DCT
    cubes[2] = np.concatenate([cubes[2, :, :, :8], cubes[2, :, :, 7::-1]], axis=2)  # mirror-symmetric in x
    cubes[3] = 255
    cubes[4] = rng.integers(126, 131, (16, 16, 16))  # tiny amplitudes: signs decided in the last bits
    for k in range(len(cubes)):
        out.append((KIND_CUBE, "cube:%02d" % k, cubes[k]))
    return out


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "refgen", "inputs.bin")
    its = items()
    with open(path, "wb") as f:
        f.write(b"VDFI" + struct.pack("<I", len(its)))
        for kind, name, a in its:
            nb = name.encode()
            f.write(struct.pack("<II", kind, len(nb)) + nb + struct.pack("<III", a.shape[2], a.shape[1], a.shape[0]))
            f.write(np.ascontiguousarray(a).tobytes())
    print(f"{len(its)} items -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
