"""CPU-only checks of the host layer above the C ABI and of the boundary itself (no compute calls):
Crop (vid_dup_finder_common/src/crop.rs:198-365, all 13 tests), MatchGroup (match_group.rs), the
(duration, path) sort and tolerance cast against the oracle, and that libvdf_b200.so loads and exports every
symbol include/vdf_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import vid_dup_finder_lib_b200 as vdf
from oracle import vdf_oracle as o
from vid_dup_finder_lib_b200 import _ffi
from vid_dup_finder_lib_b200.crop import Crop
from vid_dup_finder_lib_b200.definitions import HASH_BITS, HASH_WORDS, tolerance_to_int
from vid_dup_finder_lib_b200.video_hash import path_components, sort_order
from vid_dup_finder_lib_b200.video_hash_builder import CreationOptions, frame_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- crop.rs tests -----------------------------------------------------------------------------
@pytest.mark.parametrize("args,exp", [
    (((100, 100), 0, 0, 0, 0), (0, 0, 100, 100)),      # test_as_view_args_nocrop
    (((100, 100), 1, 0, 0, 0), (1, 0, 99, 100)),       # test_as_view_args_1pix_left
    (((100, 100), 0, 1, 0, 0), (0, 0, 99, 100)),       # test_as_view_args_1pix_right
    (((100, 100), 0, 0, 1, 0), (0, 1, 100, 99)),       # test_as_view_args_1pix_top
    (((100, 100), 0, 0, 0, 1), (0, 0, 100, 99)),       # test_as_view_args_1pix_bot
    (((100, 100), 25, 25, 25, 25), (25, 25, 50, 50)),  # test_as_view_args_four_values
    (((768, 432), 96, 96, 0, 0), (96, 0, 576, 432)),   # test_as_view_args_four_more
])
def test_as_view_args(args, exp):
    assert Crop.from_edge_offsets(*args).as_view_args() == exp


def test_from_offset_and_dims():
    assert Crop.from_topleft_and_dims((100, 100), 11, 12, 13, 14).as_view_args() == (11, 12, 13, 14)


def test_enumerate_coords_nocrop():
    c = Crop.from_edge_offsets((3, 3), 0, 0, 0, 0)
    assert len(list(c.enumerate_coords())) == 9 and len(list(c.enumerate_coords_excluded())) == 0


def test_enumerate_coords_1pixinthemiddle():
    c = Crop.from_edge_offsets((3, 3), 1, 1, 1, 1)
    assert list(c.enumerate_coords()) == [(1, 1)]
    assert sorted(c.enumerate_coords_excluded()) == sorted([(0, 0), (1, 0), (2, 0), (0, 1), (2, 1), (0, 2), (1, 2), (2, 2)])


def test_enumerate_coords_1pixinthetop():
    c = Crop.from_edge_offsets((3, 3), 1, 1, 0, 2)
    assert list(c.enumerate_coords()) == [(1, 0)]
    assert sorted(c.enumerate_coords_excluded()) == sorted([(0, 0), (2, 0), (0, 1), (1, 1), (2, 1), (0, 2), (1, 2), (2, 2)])


def test_enumerate_coords_1pixintheright():
    c = Crop.from_edge_offsets((3, 3), 2, 0, 2, 0)
    assert c == Crop.from_topleft_and_dims((3, 3), 2, 2, 1, 1)
    assert list(c.enumerate_coords()) == [(2, 2)]
    assert sorted(c.enumerate_coords_excluded()) == sorted([(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1), (0, 2), (1, 2)])


def test_crop_union_and_asserts():
    a, b = Crop.from_edge_offsets((10, 10), 1, 2, 3, 4), Crop.from_edge_offsets((10, 10), 2, 1, 0, 5)
    assert a.union(b) == Crop.from_edge_offsets((10, 10), 1, 1, 0, 4)  # crop.rs:53-68
    with pytest.raises(AssertionError):  # crop.rs:21-22
        Crop.from_edge_offsets((10, 10), 5, 5, 0, 0)


# ---- MatchGroup ---------------------------------------------------------------------------------
def test_match_group_rules():
    with pytest.raises(vdf.TooFewEntries):
        vdf.MatchGroup.new(["a"])  # match_group.rs:24 needs >= 2
    with pytest.raises(vdf.TooFewEntries):
        vdf.MatchGroup.new_with_reference("r", [])  # :40 needs >= 1
    g = vdf.MatchGroup.new(["a", "b", "c"])
    assert g.len() == 3 and g.reference() is None and list(g.contained_paths()) == ["a", "b", "c"]
    assert [list(x.duplicates()) for x in g.dup_combinations()] == [["a", "b"], ["a", "c"], ["b", "c"]]
    r = vdf.MatchGroup.new_with_reference("ref", ["a", "b"])
    assert list(r.contained_paths()) == ["a", "b", "ref"] and r.len() == 2
    assert [(x.reference(), list(x.duplicates())) for x in r.dup_combinations()] == [("ref", ["a"]), ("ref", ["b"])]


# ---- host-side pieces of the search path vs the oracle ------------------------------------------
def test_tolerance_cast_matches_oracle():
    for t in [0.0, 0.35, 0.3, 0.1, 0.0999, 1.0, 1.5, -1.0, float("nan"), 1e12, 0.3499999, 0.35000001] + [k / 1000 for k in range(0, 1001, 7)]:
        assert tolerance_to_int(t) == o.tolerance_int(t)


def test_sort_order_matches_oracle_on_awkward_paths():
    rng = np.random.default_rng(0)
    pieces = ["a", "b", "-", ".", "/", "..", "_", "0", "v"]
    paths = ["".join(rng.choice(pieces, int(rng.integers(1, 7)))) for _ in range(2000)] + ["v/%08d" % i for i in range(100)]
    dur = rng.integers(0, 4, len(paths)).astype(np.uint32)
    assert np.array_equal(sort_order(dur, paths), o.sort_order(dur, paths))
    simple = ["v/%08d" % i for i in rng.permutation(5000)] + ["/abs/x", "/abs-x", "ab/c", "ab-c", "ab"]
    d2 = rng.integers(0, 3, len(simple)).astype(np.uint32)
    assert np.array_equal(sort_order(d2, simple), o.sort_order(d2, simple))
    assert path_components("./a//b/./c/") == [(2, b""), (4, b"a"), (4, b"b"), (4, b"c")]


def test_native_sort_order_matches_the_python_rule_and_the_oracle():
    """vdf_sort_order (csrc/host.cu, multi-threaded) is Search::sort (search_algorithm.rs:55-61): stable, by
    (duration, Rust Unix Path order).  Awkward components, control bytes, prefixes, ties, and a size that takes the
    threaded merge path."""
    rng = np.random.default_rng(3)
    pieces = ["a", "b", "-", ".", "/", "..", "_", "0", "v", "\x01", " ", "ab", "~", "\u00e9"]
    paths = ["".join(rng.choice(pieces, int(rng.integers(1, 9)))) for _ in range(6000)]
    paths += ["v/%08d" % i for i in rng.permutation(300)] + ["/", "//", ".", "./", "..", "a", "a/", "a/.", "./a", "/a", "a/b", "a-b"]
    paths += paths[:500]  # exact ties: the sort must be stable
    dur = rng.integers(0, 4, len(paths)).astype(np.uint32)
    tb = vdf.HashTable(np.zeros((len(paths), 16), np.uint64), dur, paths)
    got = _ffi.sort_order(dur, *tb.path_blob())
    assert np.array_equal(got, o.sort_order(dur, paths))
    assert np.array_equal(got, sort_order(dur, paths))
    n = 300_000  # > 8192 per thread: parallel runs + merges
    big = ["v/%08d" % i for i in rng.permutation(n)]
    bd = rng.integers(590, 600, n).astype(np.uint32)
    tb = vdf.HashTable(np.zeros((n, 16), np.uint64), bd, big)
    assert np.array_equal(_ffi.sort_order(bd, *tb.path_blob()), sort_order(bd, big))
    # a library under one directory: the bytes all paths share are skipped by the sort keys; entries that ARE the shared
    # prefix, entries sharing only part of it, exact ties spread over all threads' runs
    n = 120_000
    deep = ["/data/videos/collection_%03d/clip_%08d.mp4" % (i % 500, i) for i in rng.permutation(n)]
    for extra in ([], ["/data/videos/collection_", "/data/videos/collection_001", "/data/videos"], ["/data", "/", "relative/clip.mp4"]):
        ps = deep + extra + deep[:2000]
        dd = rng.integers(598, 600, len(ps)).astype(np.uint32)
        tb = vdf.HashTable(np.zeros((len(ps), 16), np.uint64), dd, ps)
        assert np.array_equal(_ffi.sort_order(dd, *tb.path_blob()), o.sort_order(dd, ps)), extra
    assert len(_ffi.sort_order(np.zeros(0, np.uint32), *vdf.HashTable(np.zeros((0, 16), np.uint64), [], []).path_blob())) == 0


def test_match_groups_from_csr():
    paths = ["p%d" % i for i in range(6)]
    g = vdf.MatchGroup.from_csr(paths, [0, 2, 3, 6], [4, 1, 5, 0, 2, 3])
    assert [list(x.duplicates()) for x in g] == [["p4", "p1"], ["p0", "p2", "p3"]]  # rows of one entry are not groups
    r = vdf.MatchGroup.from_csr(paths, [0, 0, 2], [3, 1], references=["r0", "r1"])
    assert len(r) == 1 and r[0].reference() == "r1" and list(r[0].duplicates()) == ["p3", "p1"]


def test_constants_and_video_hash_accessors():
    assert (HASH_BITS, HASH_WORDS, vdf.DEFAULT_SEARCH_TOLERANCE, vdf.TOLERANCE_SCALING_FACTOR) == (1000, 16, 0.35, 1000.0)
    a = vdf.VideoHash.from_words([0] * 16, "a", 3)
    b = vdf.VideoHash.from_words([0xFFFFFFFFFFFFFFFF] * 16, "b", 3)
    assert a.hamming_distance(b) == 1024 and b.hamming_distance(b) == 0  # video_hash.rs:311-317 (pad bits count)
    assert a.normalized_hamming_distance(b) == 1.024 and len(b.raw_hash()) == 1000 and all(b.raw_hash())
    assert a.with_duration(9).duration == 9 and a.with_src_path("z").src_path == "z" and a < b


def test_frame_schedule_policy():  # video_hash_builder.rs:104-146
    o_ = CreationOptions()
    assert frame_schedule(1.0, o_) == ((64 * 16384, 16384), 0.0)
    assert frame_schedule(6.0, o_) == ((int(16.0 * 16384), 16384), 0.0)
    assert frame_schedule(20.0, o_) == ((int(6.4 * 16384.0), 16384), 8.0)
    assert frame_schedule(100.0, o_) == ((int(6.4 * 16384.0), 16384), 15.0)


# ---- the boundary itself ------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vdf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vdf_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_ffi.EXPORTS)
    so = os.path.join(ROOT, "vid_dup_finder_lib_b200", "libvdf_b200.so")
    assert os.path.exists(so), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(so)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    lib.vdf_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.vdf_version()


def test_integration_doc_binds_every_declared_symbol():
    """INTEGRATION.md shows the reference-side (Rust) binding of every entry point the header declares"""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vdf_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(vdf_[a-z_0-9]+)\s*\(", hdr))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    bound = set(re.findall(r"pub fn (vdf_[a-z_0-9]+)\s*\(", doc))
    assert declared <= bound, sorted(declared - bound)


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vdf.VdfError):
        vdf.Context(0)
    with pytest.raises(vdf.VdfError):
        vdf.search([vdf.VideoHash.from_words([0] * 16, "a", 1), vdf.VideoHash.from_words([0] * 16, "b", 1)], 0.35,
                   ctx=None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vid_dup_finder_lib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "vdf_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
