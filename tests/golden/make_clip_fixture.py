"""Generates tests/golden/ref_clips_gray.npz from the reference's own example clips.

Run ONLY in the build container (needs /root/reference and OpenCV's ffmpeg backend):
    python tests/golden/make_clip_fixture.py

The reference cannot run here (no cargo/rustc/ffmpeg CLI), so this does not produce reference OUTPUTS; it
captures reference INPUTS: the 16 gray frames VideoHashBuilder would feed to the compute tail
(video_hash_builder.rs:85-167: skip 15 s, fps = 64/10, take 16) for the four clips OpenCV can decode
(cat.2.mp4 and dog.2.mp4 are AV1, which this OpenCV build cannot decode).  Decoding stays on the host in the
north-star, so small decoder differences (OpenCV BGR->gray vs ffmpeg -pix_fmt gray) are out of scope; what
the fixture pins is the reference's end-to-end expectation on real content (examples/example.rs:77-82,
lib.rs:26-62): cat clips group together, dog clips group together, cats never match dogs at 0.35.
"""
import os
import sys

import cv2
import numpy as np

VIDS = "/root/reference/vid_dup_finder_lib/examples/vids"
CLIPS = ["cat.1.mp4", "cat.3.webm", "dog.1.mp4", "dog.3.webm"]
SKIP, HASH_DUR, N = 15.0, 10.0, 16


def grab(path):
    cap = cv2.VideoCapture(path)
    assert cap.isOpened(), path
    fps = cap.get(cv2.CAP_PROP_FPS)
    nfr = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    duration = nfr / fps
    assert duration >= SKIP + HASH_DUR  # the "long enough" branch, video_hash_builder.rs:138-142
    want = [int(round((SKIP + k * HASH_DUR / 64.0) * fps)) for k in range(N)]
    frames, idx = [], 0
    while len(frames) < N:
        ok, bgr = cap.read()
        assert ok
        if idx == want[len(frames)]:
            frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
        idx += 1
    return np.stack(frames), int(duration)  # duration.as_secs() as u32, video_hash_builder.rs:222


def main():
    out = {}
    durs = []
    for c in CLIPS:
        fr, d = grab(os.path.join(VIDS, c))
        out[c.replace(".", "_")] = fr
        durs.append(d)
        print(c, fr.shape, d)
    out["names"] = np.array(CLIPS)
    out["durations"] = np.array(durs, dtype=np.uint32)
    # a letterboxed + rescaled variant of cat.1 (the kind of perturbation bench/crop-*/ scripts make)
    cat = out["cat_1_mp4"]
    lb = np.full((N, 180, 320), 16, np.uint8)
    for t in range(N):
        lb[t, 18:162, 32:288] = cv2.resize(cat[t], (256, 144), interpolation=cv2.INTER_AREA)
    out["cat_1_letterboxed"] = lb
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_clips_gray.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main() if "--hashes" not in sys.argv else None


def write_oracle_hashes():
    """Second artefact: the ORACLE's hashes of those stacks (tests/golden/ref_clips_hashes.json), so a
    regression in either the oracle or the CUDA path shows up as a changed word, not just a changed group."""
    import json

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import vdf_oracle as o

    here = os.path.dirname(os.path.abspath(__file__))
    z = np.load(os.path.join(here, "ref_clips_gray.npz"))
    rec = {}
    for key in [n.replace(".", "_") for n in z["names"]] + ["cat_1_letterboxed"]:
        st, h, crop, _ = o.hash_stack(z[key], 1)
        assert st == 0
        rec[key] = {"crop_lrtb": list(crop), "hash_words_hex": [format(int(w), "016x") for w in h]}
    with open(os.path.join(here, "ref_clips_hashes.json"), "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__" and "--hashes" in sys.argv:
    write_oracle_hashes()
