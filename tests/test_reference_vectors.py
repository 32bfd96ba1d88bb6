"""Vectors produced by the REAL crates (tests/golden/refgen, run wherever cargo exists) against the oracle (CPU) and the
CUDA path (GPU).  While tests/golden/reference_vectors.json is absent these tests XFAIL with "parity unpinned": rows H3
(fast_image_resize), H4 (rustdct) and N2 (bincode) are then checked against restatements only."""
import json
import os

import numpy as np
import pytest

from oracle import vdf_oracle as o
from tests.golden import make_reference_inputs as mri

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VECTORS = os.path.join(GOLD, "reference_vectors.json")
CACHE = os.path.join(GOLD, "reference_cache.bin")
UNPINNED = ("PARITY UNPINNED: tests/golden/reference_vectors.json is missing - run tests/golden/refgen (needs cargo) to pin "
            "the resize / DCT / cache-file restatements against fast_image_resize, rustdct and bincode")


def _vectors():
    if not os.path.exists(VECTORS):
        pytest.xfail(UNPINNED)
    v = json.load(open(VECTORS))
    by_name = {it["name"]: it for it in v["items"]}
    return v, by_name


def _inputs():
    return {name: (kind, a) for kind, name, a in mri.items()}


def _bits(hexstr):
    return np.unpackbits(np.frombuffer(bytes.fromhex(hexstr), np.uint8), bitorder="little")[:1000]


def _words(item):
    return np.array([int(w, 16) for w in item["hash"]], dtype=np.uint64)


def test_kit_inputs_are_deterministic():
    """the inputs the Rust program reads are the ones this machine regenerates (numpy only)"""
    its = mri.items()
    assert len(its) == 69 and len({n for _, n, _ in its}) == 69
    assert mri.sha(its[5][2]) == mri.sha(mri.items()[5][2])
    if os.path.exists(VECTORS):
        _, by_name = _vectors()
        for kind, name, a in its:
            assert by_name[name]["input_sha256"] == mri.sha(a), name


def test_oracle_matches_the_real_crates_on_stacks():
    _, by_name = _vectors()
    for name, (kind, a) in _inputs().items():
        if kind != mri.KIND_STACK:
            continue
        ref = by_name[name]
        status, words, crop, small = o.hash_stack(a, 1)
        assert status == 0 and tuple(crop) == tuple(ref["crop"]), name
        assert small.tobytes().hex() == ref["small"], (name, "resized cube (fast_image_resize)")
        _, coefs = o.hash_from_small(small, want_coefs=True)
        assert np.array_equal((coefs[:10, :10, :10] > 0).reshape(-1).astype(np.uint8), _bits(ref["coef_positive"])), (name, "signs (rustdct)")
        assert np.array_equal(words, _words(ref)), name


def test_oracle_resize_matches_fast_image_resize_on_single_frames():
    _, by_name = _vectors()
    for name, (kind, a) in _inputs().items():
        if kind == mri.KIND_FRAME:
            assert o.resize_lanczos3(a[0]).tobytes().hex() == by_name[name]["small"], name


def test_oracle_dct_signs_match_rustdct_on_cubes():
    _, by_name = _vectors()
    for name, (kind, a) in _inputs().items():
        if kind == mri.KIND_CUBE:
            _, coefs = o.hash_from_small(a, want_coefs=True)
            assert np.array_equal((coefs[:10, :10, :10] > 0).reshape(-1).astype(np.uint8), _bits(by_name[name]["coef_positive"])), name


def test_cache_reader_matches_a_real_bincode_file(tmp_path):
    v, _ = _vectors()
    if not os.path.exists(CACHE):
        pytest.xfail(UNPINNED)
    from vid_dup_finder_lib_b200 import hash_cache as hc

    c = hc.load_hash_cache(CACHE)
    want = {e["key"]: e for e in v["cache_entries"]}
    assert len(c) == len(want)
    for i, key in enumerate(c.keys):
        e = want[key]
        assert int(c.kind[i]) == e["kind"] and [int(c.mtime_secs[i]), int(c.mtime_nanos[i])] == e["mtime"]
        if e["kind"] == 0:
            assert c.src_paths[i] == key and int(c.durations[i]) == e["duration"]
            assert np.array_equal(c.hashes[i], np.array([int(w, 16) for w in e["hash"]], dtype=np.uint64))
        elif e["kind"] == 2:
            assert c.messages[i] == e["msg"]
    out = tmp_path / "roundtrip.bin"
    hc.save_hash_cache(out, c)
    assert out.read_bytes() == open(CACHE, "rb").read()  # entries in file order: byte-identical


@pytest.mark.gpu
def test_cuda_path_matches_the_real_crates():
    _, by_name = _vectors()
    import torch

    from vid_dup_finder_lib_b200 import _ffi

    ctx = _ffi.default_context()
    for name, (kind, a) in _inputs().items():
        ref = by_name[name]
        if kind == mri.KIND_STACK:
            n_f, h, w = a.shape
            got, status, crop = ctx.hash_stacks(a.reshape(-1), _ffi.make_descs(1, w, h, n_f), 1)
            assert status[0] == 0 and tuple(int(x) for x in crop[0]) == tuple(ref["crop"]), name
            assert np.array_equal(got[0], _words(ref)), name
            d = torch.from_numpy(a).cuda()
            small = torch.empty((1, 16, 16, 16), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            ctx.hash_stacks_small_device(d.data_ptr(), _ffi.make_descs(1, w, h, n_f), 1, small.data_ptr())
            assert small.cpu().numpy().tobytes().hex() == ref["small"], name
        elif kind == mri.KIND_CUBE:
            got = ctx.hash_from_small(a[None])
            bits = np.unpackbits(got[0].view(np.uint8), bitorder="little")[:1000]
            assert np.array_equal(bits, _bits(ref["coef_positive"])), name
