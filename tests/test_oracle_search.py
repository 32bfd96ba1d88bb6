"""Pins the ORACLE's search path against every search/Hamming test the reference holds.

Each test names the reference test it re-expresses (file:line under /root/reference).  The oracle is
driven at the level the GPU boundary is (sorted hashes + durations in, index groups out); the sort is the
oracle's own restatement of Search::sort (search_algorithm.rs:55-61).
"""
import numpy as np
import pytest

from oracle import vdf_oracle as o
from tests import ref_fixtures as rf


def oracle_search(hashes, durations=None, paths=None, tol=0.35):
    """search() of video_dup_finder.rs:7-13 on the oracle -> list of groups of ORIGINAL indices."""
    n = len(hashes)
    if n == 0:
        return []
    durations = np.zeros(n, np.uint32) if durations is None else np.asarray(durations, np.uint32)
    paths = [""] * n if paths is None else paths
    order = o.sort_order(durations, paths)
    H = np.stack(hashes)[order]
    gp, mm = o.search_self(H, durations[order], o.tolerance_int(tol))
    groups = [[int(order[m]) for m in mm[gp[g]:gp[g + 1]]] for g in range(len(gp) - 1)]
    return [g for g in groups if len(g) >= 2]  # MatchGroup::new match_group.rs:21-30


def oracle_search_refs(ref_hashes, ref_dur, cand_hashes, cand_dur, tol, cand_paths=None):
    n = len(cand_hashes)
    cand_dur = np.asarray(cand_dur, np.uint32)
    order = o.sort_order(cand_dur, cand_paths or [""] * n)
    C = np.stack(cand_hashes)[order]
    rp, ci = o.search_refs(C, cand_dur[order], np.stack(ref_hashes), np.asarray(ref_dur, np.uint32), o.tolerance_int(tol))
    out = []
    for r in range(len(ref_hashes)):
        m = [int(order[k]) for k in ci[rp[r]:rp[r + 1]]]
        if m:  # video_dup_finder.rs:38-43
            out.append((r, m))
    return out


# ---- video_hash.rs:325-371 -------------------------------------------------------------------
def test_triangle_inequality():  # video_hash.rs:325-339
    rng = np.random.default_rng(1)
    for _ in range(1000):
        a, b, c = rf.random_hash(rng), rf.random_hash(rng), rf.random_hash(rng)
        assert o.hamming(a, b) <= o.hamming(a, c) + o.hamming(b, c)


def test_distance_between_two_empty_hashes_is_0():  # video_hash.rs:340-348
    assert o.hamming(rf.empty_hash(), rf.empty_hash()) == 0


def test_distance_between_two_full_hashes_is_0():  # video_hash.rs:350-357
    assert o.hamming(rf.full_hash(), rf.full_hash()) == 0


def test_symmetry():  # video_hash.rs:359-371
    rng = np.random.default_rng(2)
    for _ in range(1000):
        a, b = rf.random_hash(rng), rf.random_hash(rng)
        assert o.hamming(a, b) == o.hamming(b, a)


def test_hamming_covers_all_1024_bits():  # video_hash.rs:265-267,311-317 (SURVEY note 5)
    assert o.hamming(rf.full_hash(), rf.empty_hash()) == 1024
    rng = np.random.default_rng(3)
    for _ in range(200):
        a, b = rf.random_hash(rng), rf.random_hash(rng)
        assert o.hamming(a, b) == rf.hamming(a, b)


# ---- search_algorithm.rs:203-208 --------------------------------------------------------------
def test_searching_nothing_returns_empty_vec():
    gp, mm = o.search_self(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), o.tolerance_int(1.0))
    assert len(gp) == 1 and len(mm) == 0


# ---- tests/test_find_all.rs -------------------------------------------------------------------
def test_find_dups_finds_a_known_group():  # test_find_all.rs:137-169
    rng = np.random.default_rng(1)
    groups = rf.HashesWithDistanceSet(1, 50, 201, 100, rng)
    dups = oracle_search(groups.all_members(rng), tol=200 / 1000.0)
    assert len(dups) == 1
    assert len(dups[0]) == 50


def test_find_dups_discriminates_by_duration():  # test_find_all.rs:176-238
    rng = np.random.default_rng(2)
    groups = rf.HashesWithDistanceSet(1, 100, 201, 100, rng)
    short = groups.groups[0].members(rng)
    hashes = short + short[:50]
    dur = [50] * 100 + [250] * 50
    perm = rng.permutation(150)
    dups = oracle_search([hashes[i] for i in perm], [dur[i] for i in perm], tol=200 / 1000.0)
    dups.sort(key=len)
    assert len(dups) == 2
    assert len(dups[1]) == 100
    assert len(dups[0]) == 50
    assert all(dur[perm[i]] == 250 for i in dups[0]) and all(dur[perm[i]] == 50 for i in dups[1])


def test_find_dups_discriminates_by_distance():  # test_find_all.rs:244-269
    rng = np.random.default_rng(3)
    sets = rf.HashesWithDistanceSet(2, 100, 150, 50, rng)
    dups = oracle_search(sets.all_members(rng), tol=100 / 1000.0)
    dups.sort(key=len)
    assert len(dups) == 2
    assert len(dups[0]) == 100
    assert len(dups[1]) == 110


def test_find_with_refs():  # test_find_all.rs:273-315
    rng = np.random.default_rng(4)
    sets = rf.HashesWithDistanceSet(5, 100, 150, 50, rng)
    start = sets.groups[3].start_hash
    cands = sets.all_members(rng)
    assert len(cands) == 100 + 110 + 120 + 130 + 140
    zeros = [0] * len(cands)
    dups = oracle_search_refs([start], [0], cands, zeros, 50 / 1000.0)
    assert len(dups) == 1
    assert len(dups[0][1]) == 130
    dups2 = oracle_search_refs([sets.groups[0].start_hash, sets.groups[4].start_hash], [0, 0], cands, zeros, 50 / 1000.0)
    assert len(dups2) == 2
    assert len(dups2[0][1]) == 100  # groups come back in reference (caller) order
    assert len(dups2[1][1]) == 140


# ---- restatement cross-checks ---------------------------------------------
def _random_case(rng, n, n_clusters, max_flip):
    base = [rf.random_hash(rng) for _ in range(n_clusters)]
    hs = []
    for _ in range(n):
        h = base[int(rng.integers(0, n_clusters))].copy()
        for _ in range(int(rng.integers(0, max_flip + 1))):
            h[int(rng.integers(0, 16))] ^= np.uint64(1 << int(rng.integers(0, 64)))
        hs.append(h)
    return np.stack(hs)


@pytest.mark.parametrize("seed", range(40))
def test_literal_walk_equals_edge_list_greedy(seed):
    """search_algorithm.rs:81-171 walked literally == greedy over the (i,j)-sorted edge list, group order
    included (the form the GPU path implements)."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 300))
    H = _random_case(rng, n, int(rng.integers(1, 8)), int(rng.integers(0, 120)))
    dur = np.sort(rng.choice([0, 9, 10, 11, 12, 100, 105, 110, 111, 121, 600], n).astype(np.uint32))
    tol = int(rng.choice([0, 30, 60, 100, 150, 350, 1024]))
    gp1, m1 = o.search_self(H, dur, tol)
    e = o.self_edges(H, dur, tol)
    gp2, m2 = o.group_from_edges(n, e)
    assert np.array_equal(gp1, gp2) and np.array_equal(m1, m2)
    # group structure: members ascending, target last and smaller than every member, targets descending
    targets = [int(m1[gp1[g + 1] - 1]) for g in range(len(gp1) - 1)]
    assert targets == sorted(targets, reverse=True)
    for g in range(len(gp1) - 1):
        mem = m1[gp1[g]:gp1[g + 1] - 1]
        assert np.all(np.diff(mem.astype(np.int64)) > 0) and targets[g] < mem[0]


def test_chain_is_not_connected_components():
    """SURVEY note 2: a-b-c with a!~c yields {b,a} only."""
    a = rf.empty_hash()
    b = a.copy(); b[0] = np.uint64(0xFF)            # d(a,b)=8
    c = b.copy(); c[1] = np.uint64(0xFF)            # d(b,c)=8, d(a,c)=16
    gp, mm = o.search_self(np.stack([a, b, c]), np.zeros(3, np.uint32), 10)
    assert gp.tolist() == [0, 2] and mm.tolist() == [1, 0]


def test_duration_window_edges():
    """window end is (f64(d)*1.1) as u32 (search_algorithm.rs:99); ref window 0.95/1.05 (:174,179)."""
    assert o.self_window_thresh(10) == 11 and o.self_window_thresh(9) == 9 and o.self_window_thresh(0) == 0
    assert o.self_window_thresh(4294967295) == 4294967295  # saturating cast
    assert o.ref_window_durations(100) == (95, 105) and o.ref_window_durations(10) == (9, 10)
    h = rf.empty_hash()
    H = np.stack([h, h, h])
    gp, mm = o.search_self(H, np.array([10, 11, 12], np.uint32), 0)
    assert mm.tolist() == [1, 0]  # 12 > 11 is outside 10's window; 11 was consumed so 12 stays alone
    rp, ci = o.search_refs(H, np.array([9, 10, 11], np.uint32), H[:1], np.array([10], np.uint32), 0)
    assert ci.tolist() == [0, 1]


def test_tolerance_cast():
    for k in range(0, 1001):
        assert o.tolerance_int(k / 1000.0) == int((k / 1000.0) * 1000.0)
    assert o.tolerance_int(0.35) == 350 and o.tolerance_int(0.3) == 300
    assert o.tolerance_int(-0.5) == 0 and o.tolerance_int(float("nan")) == 0 and o.tolerance_int(1e12) == 2**32 - 1


def test_sort_is_stable_and_uses_path_components():
    dur = [5, 5, 5, 1, 5]
    paths = ["a/b", "a-b", "a/b", "z", "a"]
    order = o.sort_order(dur, paths).tolist()
    # Path::cmp: "a" < "a/b" (prefix) ; "a/b" < "a-b" because component "a" < "a-b"; equal keys keep input order
    assert order == [3, 4, 0, 2, 1]
    assert o.path_cmp("a//b/", "a/b") == 0 and o.path_cmp("a/./b", "a/b") == 0 and o.path_cmp("./a", "a") < 0
    assert o.path_cmp("/a", "a") < 0 and o.path_cmp("..", "a") < 0 and o.path_cmp("v/00000002", "v/00000010") < 0
