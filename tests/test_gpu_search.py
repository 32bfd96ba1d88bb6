"""GPU parity tests of the search path, through the C ABI (vid_dup_finder_lib_b200._ffi) and the crate-shaped API.
Bar: bit-exact edges, groups (order included) and per-reference match lists against the CPU oracle."""
import numpy as np
import pytest

import vid_dup_finder_lib_b200 as vdf
from oracle import vdf_oracle as o
from tests import ref_fixtures as rf
from tests import synth
from vid_dup_finder_lib_b200 import _ffi

pytestmark = pytest.mark.gpu
DEFAULT_VARIANT = 6  # tcgen05 cta_group::2, kind::mxf4, packed tiles (common.cuh: vdf_ctx::search_variant)


@pytest.fixture(scope="module")
def ctx():
    c = _ffi.default_context()
    c.set_shard(0, 1)
    c.set_option("search_variant", DEFAULT_VARIANT)  # everything not parametrized runs the library default
    return c


def _vh(hashes, durations=None, paths=None):
    n = len(hashes)
    durations = [0] * n if durations is None else durations
    paths = ["p/%06d" % i for i in range(n)] if paths is None else paths
    return [vdf.VideoHash.from_words(hashes[i], paths[i], int(durations[i])) for i in range(n)]


# ---- the reference's own tests, through the drop-in API ------------------------------------------
def test_searching_nothing_returns_empty_vec(ctx):  # search_algorithm.rs:203-208
    assert vdf.search([], 1.0, ctx=ctx) == []
    assert len(vdf.search([], 1.0, ctx=ctx)) == 0


def test_find_dups_finds_a_known_group(ctx):  # test_find_all.rs:137-169
    rng = np.random.default_rng(1)
    groups = rf.HashesWithDistanceSet(1, 50, 201, 100, rng)
    dups = vdf.search(_vh(groups.all_members(rng)), 200 / vdf.TOLERANCE_SCALING_FACTOR, ctx=ctx)
    assert len(dups) == 1
    assert dups[0].len() == 50


def test_find_dups_discriminates_by_duration(ctx):  # test_find_all.rs:176-238
    rng = np.random.default_rng(2)
    groups = rf.HashesWithDistanceSet(1, 100, 201, 100, rng)
    short = groups.groups[0].members(rng)
    hashes = short + short[:50]
    dur = [50] * 100 + [250] * 50
    names = ["short_%03d" % i for i in range(100)] + ["long_%03d" % i for i in range(50)]
    perm = rng.permutation(150)
    dups = vdf.search(_vh([hashes[i] for i in perm], [dur[i] for i in perm], [names[i] for i in perm]), 200 / 1000.0, ctx=ctx)
    dups = sorted(dups, key=len)
    assert len(dups) == 2
    assert dups[1].len() == 100 and all(p.startswith("short_") for p in dups[1].duplicates())
    assert dups[0].len() == 50 and all(p.startswith("long_") for p in dups[0].duplicates())


def test_find_dups_discriminates_by_distance(ctx):  # test_find_all.rs:244-269
    rng = np.random.default_rng(3)
    sets = rf.HashesWithDistanceSet(2, 100, 150, 50, rng)
    dups = vdf.search(_vh(sets.all_members(rng)), 100 / 1000.0, ctx=ctx)
    dups = sorted(dups, key=len)
    assert len(dups) == 2
    assert dups[0].len() == 100
    assert dups[1].len() == 110


def test_find_with_refs(ctx):  # test_find_all.rs:273-315
    rng = np.random.default_rng(4)
    sets = rf.HashesWithDistanceSet(5, 100, 150, 50, rng)
    cands = _vh(sets.all_members(rng))
    assert len(cands) == 100 + 110 + 120 + 130 + 140
    start = vdf.VideoHash.from_words(sets.groups[3].start_hash, "ref3", 0)
    dups = vdf.search_with_references([start], cands, 50 / 1000.0, ctx=ctx)
    assert len(dups) == 1
    assert dups[0].len() == 130 and dups[0].reference() == "ref3"
    starts = [vdf.VideoHash.from_words(sets.groups[0].start_hash, "ref0", 0),
              vdf.VideoHash.from_words(sets.groups[4].start_hash, "ref4", 0)]
    dups2 = vdf.search_with_references(starts, cands, 50 / 1000.0, ctx=ctx)
    assert len(dups2) == 2
    assert dups2[0].len() == 100 and dups2[0].reference() == "ref0"
    assert dups2[1].len() == 140 and dups2[1].reference() == "ref4"


def test_hamming_metric_properties_on_device(ctx):  # video_hash.rs:325-371 via the edge predicate
    full, empty = rf.full_hash(), rf.empty_hash()
    z = np.zeros(2, np.uint32)
    assert len(ctx.search_self(np.stack([empty, empty]), z, 0)) == 1  # d(empty, empty) = 0
    assert len(ctx.search_self(np.stack([full, full]), z, 0)) == 1    # d(full, full) = 0
    assert len(ctx.search_self(np.stack([full, empty]), z, 1023)) == 0  # all 1024 bits are compared
    assert len(ctx.search_self(np.stack([full, empty]), z, 1024)) == 1
    rng = np.random.default_rng(2)
    for _ in range(20):  # symmetry + exact distance: the edge appears exactly at tol = d, in either order
        a, b = rf.random_hash(rng), rf.random_hash(rng)
        d = o.hamming(a, b)
        for pair in (np.stack([a, b]), np.stack([b, a])):
            assert len(ctx.search_self(pair, z, d)) == 1 and len(ctx.search_self(pair, z, d - 1)) == 0


# ---- randomized parity against the oracle ---------------------------------------------------------
def _case(rng, n, n_clusters, max_flip, dur_choices):
    base = synth.random_hashes(n_clusters, seed=int(rng.integers(1 << 30)))
    pick = rng.integers(0, n_clusters, n)
    flips = rng.integers(0, 1024, (n, 1024)) < rng.integers(0, max_flip + 1, (n, 1))
    H = base[pick] ^ np.packbits(flips, axis=1, bitorder="little").view(np.uint64)
    dur = np.sort(rng.choice(dur_choices, n).astype(np.uint32))
    return np.ascontiguousarray(H), dur


SIZES = [1, 2, 3, 127, 128, 129, 255, 257, 1000, 4097]


@pytest.mark.parametrize("variant", [0, 1, 2, 5, 6], ids=["popc", "csa8x8", "csa8x4", "tcgen05_i8", "tcgen05_mxf4"])
@pytest.mark.parametrize("n", SIZES)
def test_self_search_edges_and_groups_match_oracle(ctx, n, variant):
    rng = np.random.default_rng(1000 + n)
    ctx.set_option("search_variant", variant)
    try:
        for trial in range(3):
            H, dur = _case(rng, n, max(1, n // 7), 200, [0, 9, 10, 11, 12, 100, 105, 110, 111, 121, 600] if trial else [600])
            for tol in (0, 57, 150, 350, 1024 if n <= 300 else 420):
                want_e = o.self_edges(H, dur, tol)
                got_e = ctx.search_self(H, dur, tol)
                assert np.array_equal(got_e, want_e), (n, tol, trial)
                want_gp, want_mm = o.search_self(H, dur, tol)
                got_gp, got_mm = ctx.search_self_groups(H, dur, tol)
                assert np.array_equal(got_gp, want_gp) and np.array_equal(got_mm, want_mm), (n, tol, trial)
                g2, m2 = ctx.group_greedy(n, got_e)
                assert np.array_equal(g2, want_gp) and np.array_equal(m2, want_mm)
    finally:
        ctx.set_option("search_variant", DEFAULT_VARIANT)


def test_public_api_in_caller_order_matches_oracle(ctx):
    """vdf_search / vdf_search_with_references take the caller's order and awkward paths; groups (content AND order) must
    equal the oracle's sort + greedy walk mapped back to paths (video_dup_finder.rs:7-46)."""
    rng = np.random.default_rng(77)
    n = 3000
    H, dur = _case(rng, n, 250, 180, [0, 9, 10, 11, 12, 100, 105, 110, 111, 600])
    dur = dur[rng.permutation(n)]  # caller order: unsorted durations
    pieces = ["a", "b", "-", ".", "/", "..", "_", "0", "v", "ab"]
    paths = ["".join(rng.choice(pieces, int(rng.integers(1, 7)))) + "/%05d" % i for i in range(n)]
    table = vdf.HashTable(H, dur, paths)
    order = o.sort_order(dur, paths)
    Hs, ds = np.ascontiguousarray(H[order]), dur[order]
    for tol in (0.0, 0.15, 0.35):
        tol_int = o.tolerance_int(tol)
        wgp, wmm = o.search_self(Hs, ds, tol_int)
        want = [[paths[order[int(k)]] for k in wmm[int(wgp[g]):int(wgp[g + 1])]] for g in range(len(wgp) - 1)]
        got = vdf.search(table, tol, ctx=ctx)
        assert [list(g.duplicates()) for g in got] == [w for w in want if len(w) >= 2]
        assert all(g.reference() is None for g in got)
    nr = 200
    R = H[rng.integers(0, n, nr)] ^ np.packbits(rng.integers(0, 1024, (nr, 1024)) < 30, axis=1, bitorder="little").view(np.uint64)
    rdur = rng.choice([9, 10, 100, 105, 600], nr).astype(np.uint32)
    rpaths = ["ref/%d" % i for i in range(nr)]
    rp, ci = o.search_refs(Hs, ds, R, rdur, 200)
    want = [(rpaths[r], [paths[order[int(k)]] for k in ci[int(rp[r]):int(rp[r + 1])]]) for r in range(nr) if rp[r + 1] > rp[r]]
    got = vdf.search_with_references(vdf.HashTable(R, rdur, rpaths), table, 0.2, ctx=ctx)
    assert [(g.reference(), list(g.duplicates())) for g in got] == want
    assert sum(ctx.last_phases()) > 0


def test_stage_sorted_leaves_the_sorted_table_resident(ctx):
    """vdf_stage_sorted = Search::seed + sort with the table left in HBM: the permutation is the oracle's, and the device
    pointers it returns feed vdf_search_self_device (what every rank of a multi-GPU search does) - also when the upload goes
    into caller-owned buffers"""
    import torch

    rng = np.random.default_rng(78)
    n = 2500
    H, dur = _case(rng, n, 200, 180, [0, 9, 10, 11, 12, 100, 105, 110, 111, 600])
    dur = dur[rng.permutation(n)]
    pieces = ["a", "b", "-", ".", "/", "..", "_", "0", "v", "ab"]
    paths = ["".join(rng.choice(pieces, int(rng.integers(1, 7)))) + "/%05d" % i for i in range(n)]
    table = vdf.HashTable(H, dur, paths)
    want_order = o.sort_order(dur, paths)
    want_keys = o.self_edges(np.ascontiguousarray(H[want_order]), dur[want_order], 300)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    with torch.cuda.stream(stream):
        keys = torch.empty(1 << 16, dtype=torch.int64, device=dev)
        own_h = torch.empty((n, 16), dtype=torch.int64, device=dev)
        own_d = torch.empty(n, dtype=torch.int32, device=dev)
        for dst in ((0, 0), (own_h.data_ptr(), own_d.data_ptr())):
            order, p_hash, p_dur = ctx.stage_sorted(table.hashes, table.durations, *table.path_blob(), d_hash_dst=dst[0], d_dur_dst=dst[1])
            assert np.array_equal(order, want_order)
            if dst[0]:
                assert (p_hash, p_dur) == dst
            torch.cuda.current_stream().synchronize()
            cnt = ctx.search_self_device(p_hash, p_dur, n, 300, keys.data_ptr(), keys.numel())
            got = keys[:cnt].cpu().numpy().view(np.uint64)
            assert np.array_equal(np.stack([got >> np.uint64(32), got & np.uint64(0xFFFFFFFF)], axis=1), want_keys)
        assert np.array_equal(own_d.cpu().numpy().view(np.uint32), dur[want_order])
        assert np.array_equal(own_h.cpu().numpy().view(np.uint64), H[want_order])


def test_peer_exchange_needs_its_set_up(ctx):
    """the fused exchange refuses to run half configured (its real test needs >= 2 GPUs: scripts/dist_check.py)"""
    H, dur = _case(np.random.default_rng(5), 300, 30, 100, [600])
    with pytest.raises(_ffi.VdfError):
        ctx.peer_open(0, 2, b"\0" * 128)  # no vdf_peer_alloc before
    ctx.set_option("exchange", 1)
    try:
        with pytest.raises(_ffi.VdfError):
            ctx.search_self(H, dur, 300)  # "exchange" on, no peers mapped
    finally:
        ctx.set_option("exchange", 0)
    assert len(ctx.peer_alloc(1000)) == 64  # allocation and IPC export work on one GPU; world >= 2 is required to open
    with pytest.raises(_ffi.VdfError):
        ctx.peer_open(0, 1, b"\0" * 64)
    ctx.peer_close()
    assert np.array_equal(ctx.search_self(H, dur, 300), o.self_edges(H, dur, 300))


def test_greedy_rule_is_not_connected_components(ctx):
    a = rf.empty_hash()
    b = a.copy(); b[0] = np.uint64(0xFF)
    c = b.copy(); c[1] = np.uint64(0xFF)
    gp, mm = ctx.search_self_groups(np.stack([a, b, c]), np.zeros(3, np.uint32), 10)
    assert gp.tolist() == [0, 2] and mm.tolist() == [1, 0]  # a-b-c with a!~c: {b, a}; c stays alone


def test_components_mode_is_a_gpu_union_find(ctx):
    """Optional grouping mode (the north-star's "GPU union-find"): connected components of the edge
    graph, checked against the sequential DisjointSet walk (disjoint_set.rs:22-44) and against its own tests' cases
    (disjoint_set.rs:217-335: one set, an extra item, two sets, contains_pair)."""
    gp, mm = ctx.group_components(20, [[1, 2]])
    assert gp.tolist() == [0, 2] and mm.tolist() == [2, 1]
    gp, mm = ctx.group_components(20, [[1, 2], [2, 3]])  # test_insert_extra_item_to_single_set
    assert gp.tolist() == [0, 3] and mm.tolist() == [2, 3, 1]
    gp, mm = ctx.group_components(20, [[1, 2], [11, 12], [1, 3]])  # test_insert_two_sets / test_contains_pair
    assert gp.tolist() == [0, 2, 5] and mm.tolist() == [12, 11, 2, 3, 1]
    # a chain is ONE component but only {a, b} under the reference's rule
    chain = [[0, 1], [1, 2]]
    assert ctx.group_components(3, chain)[1].tolist() == [1, 2, 0] and ctx.group_greedy(3, chain)[1].tolist() == [1, 0]
    rng = np.random.default_rng(21)
    for n, ne in ((50, 30), (2000, 1500), (200000, 150000), (5000, 60000)):
        e = rng.integers(0, n, (ne, 2)).astype(np.uint64)
        e = np.unique(np.sort(e[e[:, 0] != e[:, 1]], axis=1), axis=0)
        wgp, wmm = o.group_components(n, e)
        ggp, gmm = ctx.group_components(n, e)
        assert np.array_equal(ggp, wgp) and np.array_equal(gmm, wmm), (n, ne)
    # through the public entry point: ctx option "grouping" = 1; star-shaped clusters come out identical in both modes
    H, dur = _case(rng, 3000, 400, 60, [600])
    paths = ["p/%05d" % i for i in range(len(dur))]
    table = vdf.HashTable(H, dur, paths)
    greedy = vdf.search(table, 0.35, ctx=ctx)
    ctx.set_option("grouping", 1)
    try:
        comps = vdf.search(table, 0.35, ctx=ctx)
        e = ctx.search_self(H, dur, 350)
        wgp, wmm = o.group_components(len(dur), e)
        assert [list(g.duplicates()) for g in comps] == [[paths[int(k)] for k in wmm[int(wgp[g]):int(wgp[g + 1])]] for g in range(len(wgp) - 1)]
        assert sum(g.len() for g in comps) >= sum(g.len() for g in greedy)
    finally:
        ctx.set_option("grouping", 0)


def test_long_dependency_chain(ctx):
    """a path graph 0-1-2-...-(L-1) needs L rounds of the parallel greedy rule: targets are the even vertices."""
    L = 600
    edges = np.stack([np.arange(L - 1), np.arange(1, L)], axis=1).astype(np.uint64)
    gp, mm = ctx.group_greedy(L, edges)
    want_gp, want_mm = o.group_from_edges(L, edges)
    assert np.array_equal(gp, want_gp) and np.array_equal(mm, want_mm)
    assert len(gp) - 1 == L // 2


def test_random_edge_lists_group_like_the_oracle(ctx):
    rng = np.random.default_rng(77)
    for n, m in [(10, 12), (200, 150), (200, 2000), (5000, 20000), (1000, 1)]:
        i = rng.integers(0, n, m)
        j = rng.integers(0, n, m)
        e = np.unique(np.stack([np.minimum(i, j), np.maximum(i, j)], axis=1)[i != j], axis=0).astype(np.uint64)
        gp, mm = ctx.group_greedy(n, e)
        want_gp, want_mm = o.group_from_edges(n, e)
        assert np.array_equal(gp, want_gp) and np.array_equal(mm, want_mm)


@pytest.mark.parametrize("variant", [0, 5, 6], ids=["popc", "tcgen05_i8", "tcgen05_mxf4"])
@pytest.mark.parametrize("n_cand,n_ref", [(1, 1), (300, 5), (129, 257), (5000, 700)])
def test_ref_search_matches_oracle(ctx, n_cand, n_ref, variant):
    ctx.set_option("search_variant", variant)
    try:
        _ref_search_case(ctx, n_cand, n_ref)
    finally:
        ctx.set_option("search_variant", DEFAULT_VARIANT)


def _ref_search_case(ctx, n_cand, n_ref):
    rng = np.random.default_rng(n_cand * 31 + n_ref)
    durs = [0, 9, 10, 11, 95, 100, 105, 106, 600, 630, 631]
    C, cdur = _case(rng, n_cand, max(1, n_cand // 9), 150, durs)
    R = C[rng.integers(0, n_cand, n_ref)] ^ np.packbits(rng.integers(0, 1024, (n_ref, 1024)) < 40, axis=1, bitorder="little").view(np.uint64)
    rdur = rng.choice(durs, n_ref).astype(np.uint32)  # caller order: NOT sorted
    for tol in (0, 60, 200, 350):
        want_rp, want_ci = o.search_refs(C, cdur, R, rdur, tol)
        got_rp, got_ci = ctx.search_refs(C, cdur, R, rdur, tol)
        assert np.array_equal(got_rp, want_rp) and np.array_equal(got_ci, want_ci), tol


def test_saturated_tolerance_matches_everything_in_the_window(ctx):
    """(tolerance * 1000.0) as u32 saturates (search_algorithm.rs:82): a huge tolerance must behave like 1024 on every kernel"""
    rng = np.random.default_rng(8)
    H, dur = _case(rng, 700, 100, 300, [10, 11, 100])
    want = o.self_edges(H, dur, 0xFFFFFFFF)
    assert len(want) == o.self_window_pairs(dur)
    for variant in (0, 2, 5, 6):
        ctx.set_option("search_variant", variant)
        try:
            for tol in (1024, 1025, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF):
                assert np.array_equal(ctx.search_self(H, dur, tol), want), (variant, tol)
        finally:
            ctx.set_option("search_variant", DEFAULT_VARIANT)
    g = vdf.search(vdf.HashTable(H, dur, ["p/%04d" % i for i in range(700)]), 1e12, ctx=ctx)
    assert sum(x.len() for x in g) > 0


def test_empty_and_degenerate_inputs(ctx):
    e = ctx.search_self(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), 350)
    assert e.shape == (0, 2)
    gp, mm = ctx.search_self_groups(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), 350)
    assert gp.tolist() == [0] and len(mm) == 0
    rp, ci = ctx.search_refs(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), synth.random_hashes(3), np.zeros(3, np.uint32), 350)
    assert rp.tolist() == [0, 0, 0, 0] and len(ci) == 0
    rp, ci = ctx.search_refs(synth.random_hashes(3), np.zeros(3, np.uint32), np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), 350)
    assert rp.tolist() == [0]
    # maximum durations: the 1.1x window saturates at u32::MAX (Rust `as u32`)
    h = rf.empty_hash()
    dur = np.array([4294967295 - 5, 4294967295], np.uint32)
    assert ctx.search_self(np.stack([h, h]), dur, 0).tolist() == [[0, 1]]


def test_edge_buffer_grows_and_caps(ctx):
    H = np.zeros((600, 16), np.uint64)  # all identical: 600*599/2 = 179700 edges
    dur = np.zeros(600, np.uint32)
    ctx.set_option("initial_edges", 1000)
    try:
        e = ctx.search_self(H, dur, 0)
        assert len(e) == 600 * 599 // 2
        gp, mm = ctx.search_self_groups(H, dur, 0)
        assert gp.tolist() == [0, 600] and mm.tolist() == list(range(1, 600)) + [0]
        ctx.set_option("max_edges", 5000)
        with pytest.raises(vdf.VdfError) as ei:
            ctx.search_self(H, dur, 0)
        assert ei.value.code == _ffi.ERR_EDGE_OVERFLOW
    finally:
        ctx.set_option("initial_edges", 1 << 22)
        ctx.set_option("max_edges", 1 << 28)


@pytest.mark.parametrize("variant", [0, 2, 5, 6], ids=["popc", "csa8x4", "tcgen05_i8", "tcgen05_mxf4"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_partition_the_pair_matrix(ctx, world, variant):
    rng = np.random.default_rng(5)
    H, dur = _case(rng, 3000, 300, 200, [600, 610, 650, 700])
    want = o.self_edges(H, dur, 300)
    parts = []
    ctx.set_option("search_variant", variant)
    try:
        for r in range(world):
            ctx.set_shard(r, world)
            parts.append(ctx.search_self(H, dur, 300))
    finally:
        ctx.set_shard(0, 1)
        ctx.set_option("search_variant", DEFAULT_VARIANT)
    assert sum(len(p) for p in parts) == len(want)  # disjoint
    allp = np.concatenate(parts)
    allp = allp[np.lexsort((allp[:, 1], allp[:, 0]))]
    assert np.array_equal(allp, want)


@pytest.mark.parametrize("order,a_tmem", [(0, 0), (1, 0), (1, 1)])
def test_variant6_kernel_options_do_not_change_results(ctx, order, a_tmem):
    """work-unit order (chunk-major / row-pair-major) and row operand placement (tensor memory / shared memory) of the
    kind::mxf4 kernel, incl. a sharded run with small chunks so that the unit list has many chunks per row pair"""
    rng = np.random.default_rng(77)
    H, dur = _case(rng, 5000, 400, 200, [600, 610, 650, 700])
    want = o.self_edges(H, dur, 320)
    ctx.set_option("search_variant", 6)
    ctx.set_option("tc_unit_order", order)
    ctx.set_option("tc_a_tmem", a_tmem)
    ctx.set_option("tc_chunk", 2)
    try:
        assert np.array_equal(ctx.search_self(H, dur, 320), want)
        parts = []
        for r in range(3):
            ctx.set_shard(r, 3)
            parts.append(ctx.search_self(H, dur, 320))
        allp = np.concatenate(parts)
        assert np.array_equal(allp[np.lexsort((allp[:, 1], allp[:, 0]))], want)
    finally:
        ctx.set_shard(0, 1)
        ctx.set_option("tc_chunk", 0)
        ctx.set_option("tc_a_tmem", 1)
        ctx.set_option("tc_unit_order", 0)
        ctx.set_option("search_variant", DEFAULT_VARIANT)


# ---- round 2: the fold (column popcounts inside the contraction), prepared tables, GPU-side sort, multi-device contexts ----
def _real_case(rng, n, n_clusters, max_flip, dur_choices, dense=False):
    """like _case, but bits 1000..1023 stay zero as in every real VideoHash: these tables take the fold path"""
    H, dur = _case(rng, n, n_clusters, max_flip, dur_choices)
    if dense:  # popcounts over the whole range 0..1000, not just ~500: all-ones, all-zero and biased hashes
        k = rng.integers(0, 1025, (n, 1))
        H = np.packbits(rng.integers(0, 1024, (n, 1024)) < k, axis=1, bitorder="little").view(np.uint64).copy()
        H[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
        H[1] = 0
        H[2] = H[0]
    H[:, 15] &= np.uint64((1 << 40) - 1)
    return np.ascontiguousarray(H), dur


@pytest.mark.parametrize("dense", [False, True], ids=["clusters", "dense"])
@pytest.mark.parametrize("a_tmem", [1, 0])
def test_fold_is_exact_at_every_tolerance(ctx, dense, a_tmem):
    """variant 6 with acc = 2 dot - pc(j) + C (tc_fold, the default for real hashes) against the oracle and against the
    popcount-screen epilogue (tc_fold 0), from tolerance 0 through the tolerances where round 1's screen gave up (0.40 ..
    0.45), to everything-matches; dense hashes drive pc(j) through all of 0..1000 (every digit string of the fold unit)."""
    rng = np.random.default_rng(4242 + dense)
    ctx.set_option("tc_a_tmem", a_tmem)
    try:
        for n in (130, 1500):
            H, dur = _real_case(rng, n, max(1, n // 9), 260, [600, 610, 650], dense=dense)
            for tol in (0, 57, 350, 400, 420, 450, 500, 799, 800, 801, 1000, 1024, 0xFFFFFFFF):
                want = o.self_edges(H, dur, tol)
                ctx.set_option("tc_fold", -1)
                got = ctx.search_self(H, dur, tol)
                assert np.array_equal(got, want), (n, tol, "fold")
                ctx.set_option("tc_fold", 0)
                assert np.array_equal(ctx.search_self(H, dur, tol), want), (n, tol, "no fold")
        # reference search: the fold needs zero pad bits in BOTH operands
        C, cdur = _real_case(rng, 2000, 150, 200, [95, 100, 105, 600], dense=dense)
        R, rdur = C[rng.integers(0, 2000, 300)].copy(), rng.choice([100, 600], 300).astype(np.uint32)
        R[:, 3] ^= np.uint64(0xFFFF)
        for fold in (-1, 0):
            ctx.set_option("tc_fold", fold)
            for tol in (30, 350, 430):
                wrp, wci = o.search_refs(C, cdur, R, rdur, tol)
                grp, gci = ctx.search_refs(C, cdur, R, rdur, tol)
                assert np.array_equal(grp, wrp) and np.array_equal(gci, wci), (fold, tol)
        R[7, 15] |= np.uint64(1 << 63)  # one reference with a pad bit: the call falls back, results unchanged in kind
        wrp, wci = o.search_refs(C, cdur, R, rdur, 350)
        ctx.set_option("tc_fold", -1)
        grp, gci = ctx.search_refs(C, cdur, R, rdur, 350)
        assert np.array_equal(grp, wrp) and np.array_equal(gci, wci)
    finally:
        ctx.set_option("tc_fold", -1)
        ctx.set_option("tc_a_tmem", 1)


def test_many_matches_per_warp(ctx):
    """tolerances where a large share of all pairs match: the warp-aggregated append (one atomic per warp and 64 columns)
    and the count-then-grow edge buffer"""
    rng = np.random.default_rng(99)
    H, dur = _real_case(rng, 3000, 40, 120, [600])
    ctx.set_option("initial_edges", 4096)
    try:
        for tol in (200, 480, 520):
            want = o.self_edges(H, dur, tol)
            assert len(want) > 50_000
            assert np.array_equal(ctx.search_self(H, dur, tol), want), tol
            gp, mm = ctx.search_self_groups(H, dur, tol)
            wgp, wmm = o.search_self(H, dur, tol)
            assert np.array_equal(gp, wgp) and np.array_equal(mm, wmm)
    finally:
        ctx.set_option("initial_edges", 1 << 22)


def test_prepared_table_serves_many_searches(ctx):
    """vdf_table = `Search::from` once: tolerance sweeps, shard changes, kernel changes and reference searches on the same
    resident table, each equal to the oracle"""
    import torch

    rng = np.random.default_rng(31)
    H, dur = _real_case(rng, 5000, 400, 200, [9, 10, 11, 100, 105, 110, 600])
    t = ctx.table_create(H, dur)
    assert len(t) == 5000
    dev = torch.device("cuda", ctx.device)
    keys = torch.empty(1 << 20, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def edges(cnt):
        k = keys[:cnt].cpu().numpy().view(np.uint64)
        return np.stack([k >> np.uint64(32), k & np.uint64(0xFFFFFFFF)], axis=1)

    try:
        for tol in (0, 150, 350, 420):
            assert np.array_equal(edges(t.search_self_device(tol, keys.data_ptr(), keys.numel())), o.self_edges(H, dur, tol))
            gp, mm = t.search_self_groups(tol)
            wgp, wmm = o.search_self(H, dur, tol)
            assert np.array_equal(gp, wgp) and np.array_equal(mm, wmm)
        want = o.self_edges(H, dur, 300)
        parts = []
        for r in range(3):  # the work-unit plan follows the shard
            ctx.set_shard(r, 3)
            parts.append(edges(t.search_self_device(300, keys.data_ptr(), keys.numel())))
        ctx.set_shard(0, 1)
        allp = np.concatenate(parts)
        assert np.array_equal(allp[np.lexsort((allp[:, 1], allp[:, 0]))], want)
        for variant in (0, 5, 6):  # the table re-packs itself for another kernel
            ctx.set_option("search_variant", variant)
            assert np.array_equal(edges(t.search_self_device(300, keys.data_ptr(), keys.numel())), want), variant
        ctx.set_option("search_variant", DEFAULT_VARIANT)
        R = H[rng.integers(0, 5000, 200)].copy()
        R[:, 5] ^= np.uint64(0xFF)
        rdur = rng.choice([10, 100, 600], 200).astype(np.uint32)
        d_r = torch.from_numpy(R.view(np.int64)).to(dev)
        d_rd = torch.from_numpy(rdur.view(np.int32)).to(dev)
        torch.cuda.synchronize()
        for tol in (100, 350):
            cnt = t.search_refs_device(0, d_r.data_ptr(), d_rd.data_ptr(), 200, tol, keys.data_ptr(), keys.numel())
            wrp, wci = o.search_refs(H, dur, R, rdur, tol)
            k = keys[:cnt].cpu().numpy().view(np.uint64)
            rows = np.repeat(np.arange(200, dtype=np.uint64), np.diff(wrp).astype(np.int64))
            assert np.array_equal(k, (rows << np.uint64(32)) | wci)
    finally:
        ctx.set_shard(0, 1)
        ctx.set_option("search_variant", DEFAULT_VARIANT)
        t.close()
    empty = ctx.table_create(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32))
    assert empty.search_self_groups(350)[0].tolist() == [0]
    empty.close()


def test_sort_ties_are_finished_on_the_host(ctx):
    """the GPU sorts (duration, 16 key bytes, index); paths that still agree after their first 16 distinguishing bytes are
    ordered inside their runs by the full Rust Path comparison -- the result is Search::sort's permutation either way"""
    rng = np.random.default_rng(123)
    n = 4000
    H, dur = _case(rng, n, 300, 150, [7, 600])
    dur = dur[rng.permutation(n)]
    deep = ["/srv/media/library/shows/a_very_long_series_name_%d/season_%02d/" % (i % 3, i % 5) for i in range(n)]
    names = ["episode_%05d.mkv" % int(v) for v in rng.permutation(n)]
    paths = [d + f for d, f in zip(deep, names)]
    paths[10] = paths[11]  # identical paths: stability decides
    paths[12] = paths[11] + "/"  # trailing separator: the same components as paths[11]
    table = vdf.HashTable(H, dur, paths)
    want_order = o.sort_order(dur, paths)
    order, _, _ = ctx.stage_sorted(table.hashes, table.durations, *table.path_blob())
    assert np.array_equal(order, want_order)
    Hs, ds = np.ascontiguousarray(H[want_order]), dur[want_order]
    wgp, wmm = o.search_self(Hs, ds, 300)
    want = [[paths[want_order[int(k)]] for k in wmm[int(wgp[g]):int(wgp[g + 1])]] for g in range(len(wgp) - 1)]
    got = vdf.search(table, 0.3, ctx=ctx)
    assert [list(g.duplicates()) for g in got] == want
    # and a table with one entry, and one whose paths are all equal
    one = vdf.HashTable(H[:1], dur[:1], ["x"])
    assert len(vdf.search(one, 0.3, ctx=ctx)) == 0
    same = vdf.HashTable(np.repeat(H[:1], 5, axis=0), np.zeros(5, np.uint32), ["same/path"] * 5)
    g = vdf.search(same, 0.0, ctx=ctx)
    assert len(g) == 1 and g[0].len() == 5


def test_two_contexts_in_one_process(ctx):
    """cudaFuncSetAttribute is per device and the opt-in flags live in the context: a second context, on
    another GPU when the box has one, launches every >48 KB kernel"""
    import torch

    dev = 1 if torch.cuda.device_count() > 1 else 0
    other = _ffi.Context(dev)
    try:
        rng = np.random.default_rng(5)
        H, dur = _case(rng, 700, 60, 150, [600])
        want = o.self_edges(H, dur, 300)
        for variant in (0, 2, 5, 6):
            other.set_option("search_variant", variant)
            assert np.array_equal(other.search_self(H, dur, 300), want)
            ctx.set_option("search_variant", variant)
            assert np.array_equal(ctx.search_self(H, dur, 300), want)
        # the hashing kernels' opt-in (224 KB of shared memory for the fused kernel) is per device too
        st = synth.frame_stacks(5, 320, 180, seed=12).numpy()
        st[1, :, :25, :] = 16
        descs = _ffi.make_descs(5, 320, 180)
        a, b = other.hash_stacks(st.reshape(-1), descs, 1), ctx.hash_stacks(st.reshape(-1), descs, 1)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[2][1][2] == 25
    finally:
        ctx.set_option("search_variant", DEFAULT_VARIANT)
        other.close()


def test_multi_device_context_matches_one_device(ctx):
    """vdf_ctx_create_multi: one context, several GPUs, one host thread per device, results identical to one device"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(8)
    n = 20000
    H, dur = _real_case(rng, n, 1500, 200, [9, 10, 11, 100, 105, 110, 600])
    dur = dur[rng.permutation(n)]
    paths = ["lib/%02d/v%05d" % (i % 7, i) for i in range(n)]
    table = vdf.HashTable(H, dur, paths)
    m = _ffi.Context(list(range(min(torch.cuda.device_count(), 8))))
    try:
        assert m.device_count >= 2
        for tol in (0.1, 0.35):
            want = vdf.search(table, tol, ctx=ctx)
            got = vdf.search(table, tol, ctx=m)
            assert got == want and len(got) > 0
        m.set_option("initial_edges", 2048)  # overflow: every device re-allocates together
        assert vdf.search(table, 0.35, ctx=m) == vdf.search(table, 0.35, ctx=ctx)
        R = vdf.HashTable(H[::40] ^ np.uint64(1), dur[::40], ["r/%d" % i for i in range(len(H[::40]))])
        assert vdf.search_with_references(R, table, 0.3, ctx=m) == vdf.search_with_references(R, table, 0.3, ctx=ctx)
        few = vdf.HashTable(H[:3], dur[:3], paths[:3])  # fewer candidates than devices: empty slices still reach the barrier
        assert vdf.search_with_references(R, few, 0.3, ctx=m) == vdf.search_with_references(R, few, 0.3, ctx=ctx)
        # vdf_hash_stacks on the multi-device context: stacks sharded over the devices, one fused launch each
        st = synth.frame_stacks(37, 320, 180, seed=13).numpy()
        st[::4, :, :, :30] = 20
        descs = _ffi.make_descs(37, 320, 180)
        a, b = m.hash_stacks(st.reshape(-1), descs, 1), ctx.hash_stacks(st.reshape(-1), descs, 1)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and b[2][0][0] == 30
    finally:
        m.close()


def test_parity_at_65536_planted(ctx):
    """full CPU-vs-GPU parity at N = 2^16 on the planted-duplicate generator (log-normal durations
    exercise the windows)."""
    n = 1 << 16
    H, _ = synth.planted_hashes(n, seed=synth.SEED)
    dur = synth.lognormal_durations(n)
    paths = synth.paths(n)
    order = o.sort_order(dur, paths)
    Hs, ds = np.ascontiguousarray(H[order]), dur[order]
    for tol in (100, 350):
        want = o.self_edges(Hs, ds, tol)
        got = ctx.search_self(Hs, ds, tol)
        assert np.array_equal(got, want)
        wgp, wmm = o.search_self(Hs, ds, tol)
        ggp, gmm = ctx.search_self_groups(Hs, ds, tol)
        assert np.array_equal(ggp, wgp) and np.array_equal(gmm, wmm)
    assert ctx.self_window_pairs(ds) == o.self_window_pairs(ds)


def test_full_size_1m_properties(ctx):
    """BASELINE config: all-pairs on 1 M synthetic hashes (equal durations).  The oracle cannot finish 5e11
    pairs, so: (1) every planted duplicate within tolerance of its source is found, (2) the full edge rows of
    256 sampled entries equal the oracle's brute force over all 1 M candidates, (3) groups follow from the
    edges by the oracle's greedy rule, (4) edges are sorted and i < j."""
    n = 1_000_000
    tol = 350
    H, src = synth.planted_hashes(n, seed=synth.SEED)
    dur = np.full(n, 600, np.uint32)
    e = ctx.search_self(H, dur, tol)
    assert np.all(e[:, 0] < e[:, 1])
    keys = (e[:, 0] << np.uint64(32)) | e[:, 1]
    assert np.all(np.diff(keys.astype(np.int64)) > 0)
    have = set(keys.tolist())
    # (1) planted pairs
    dups = np.nonzero(src != np.arange(n, dtype=np.uint64))[0]
    d = rf.popcount64(H[dups] ^ H[src[dups].astype(np.int64)])
    close = d <= tol
    assert close.sum() > 0.8 * len(dups)  # ~10 % of the sources are perturbed copies themselves
    for i, s in zip(dups[close][:20000], src[dups][close][:20000]):
        a, b = (int(s), int(i)) if s < i else (int(i), int(s))
        assert ((a << 32) | b) in have
    # (2) sampled rows vs oracle brute force
    rng = np.random.default_rng(9)
    rows = np.unique(np.concatenate([rng.integers(0, n, 192), dups[:64]]))
    rp, ci = o.search_refs(H, dur, H[rows], dur[rows], tol)
    for k, r in enumerate(rows):
        want = set(int(c) for c in ci[rp[k]:rp[k + 1]] if c != r)
        got = set(e[e[:, 0] == r][:, 1].tolist()) | set(e[e[:, 1] == r][:, 0].tolist())
        assert got == want, r
    # (3) grouping
    gp, mm = ctx.group_greedy(n, e)
    wgp, wmm = o.group_from_edges(n, e)
    assert np.array_equal(gp, wgp) and np.array_equal(mm, wmm)
