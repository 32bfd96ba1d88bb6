"""world_size-2 (and 3) gloo tests of the multi-GPU host plumbing (vid_dup_finder_lib_b200/dist.py) on CPU: the
variable-length edge all-gather + merge that follows the sharded pair-matrix kernel, the candidate-slice
arithmetic of the sharded reference search, and the CSR assembly.  Edge lists come from the oracle, split
the way ranks would hold them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vdf_oracle as o
from tests import synth
from vid_dup_finder_lib_b200 import dist as vdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 3000
        H, _ = synth.planted_hashes(n, seed=5, dup_frac_den=3, max_flip=200)
        dur = np.sort(synth.lognormal_durations(n, seed=5))
        tol = 300
        edges = o.self_edges(H, dur, tol)
        keys = (edges[:, 0] << np.uint64(32)) | edges[:, 1]
        # rank r holds the edges of the (row tile + col tile) units it owns, in arbitrary (unsorted) order
        unit_owner = ((edges[:, 0] // 128) + (edges[:, 1] // 128 - edges[:, 0] // 128) // 4) % world
        rng = np.random.default_rng(rank)
        mine = rng.permutation(keys[unit_owner == rank])
        merged = vdist.merge_keys(torch.from_numpy(mine.view(np.int64)))
        ok_self = np.array_equal(merged.numpy().view(np.uint64), keys)
        gp, mm = o.group_from_edges(n, np.stack([merged.numpy().view(np.uint64) >> np.uint64(32),
                                                 merged.numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)], axis=1))
        wgp, wmm = o.search_self(H, dur, tol)
        ok_groups = np.array_equal(gp, wgp) and np.array_equal(mm, wmm)

        # reference search: candidate slices per rank, keys carry global candidate indices
        R = H[::37][:50]
        rdur = dur[::37][:50]
        b, e = vdist.shard_range(n, rank, world)
        rp, ci = o.search_refs(H[b:e], dur[b:e], R, rdur, tol)
        rows = np.repeat(np.arange(len(R), dtype=np.uint64), np.diff(rp).astype(np.int64))
        local = (rows << np.uint64(32)) | (ci + np.uint64(b))
        allk = vdist.merge_keys(torch.from_numpy(local.view(np.int64)))
        grp, gci = vdist.csr_from_keys(allk.numpy(), len(R))
        wrp, wci = o.search_refs(H, dur, R, rdur, tol)
        ok_refs = np.array_equal(grp, wrp) and np.array_equal(gci, wci)

        # empty contribution from one rank and empty everywhere
        part = torch.arange(5, dtype=torch.int64) if rank == 0 else torch.zeros(0, dtype=torch.int64)
        ok_empty = vdist.allgather_varlen(part).tolist() == [0, 1, 2, 3, 4]
        ok_empty &= vdist.allgather_varlen(torch.zeros(0, dtype=torch.int64)).numel() == 0
        q.put((rank, ok_self, ok_groups, ok_refs, ok_empty))
    finally:
        dist.destroy_process_group()


class _StubCtx:
    """stands in for _ffi.Context in the collective logic of the fused exchange (set-up, overflow -> all ranks re-allocate
    together -> retry); the real thing needs >= 2 GPUs and is checked by scripts/dist_check.py"""

    def __init__(self, rank):
        self.rank, self.peer_capacity, self.peer_world, self.log = rank, 0, 0, []

    def peer_alloc(self, cap):
        self.peer_capacity, self.peer_world = int(cap), 0
        self.log.append(("alloc", int(cap)))
        return bytes([self.rank]) * 64

    def peer_open(self, rank, world, handles):
        assert rank == self.rank and handles == b"".join(bytes([r]) * 64 for r in range(world))  # rank order
        self.peer_world = world
        self.log.append(("open", world))

    def peer_close(self):
        self.peer_capacity, self.peer_world = 0, 0
        self.log.append(("close",))

    def set_option(self, k, v):
        self.log.append((k, v))


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _StubCtx(rank)
        assert vdist.enable_peer_exchange(ctx, capacity=100)
        total = 700  # matches over all ranks: the same count reaches every rank, as from the library

        def fn(ptr, cap):
            return total if cap >= total else -total

        keys = vdist._run_exchange(ctx, fn, torch.device("cpu"), None)
        ok = keys.numel() == total and ctx.peer_world == world and ctx.peer_capacity >= total
        ok &= [e for e in ctx.log if e[0] in ("alloc", "close")] == [("alloc", 100), ("close",), ("alloc", ctx.peer_capacity)]
        ok &= ctx.log.count(("exchange", 1)) == 2 and ctx.log[-1] == ("exchange", 0)
        vdist.disable_peer_exchange(ctx)
        ok &= ctx.peer_world == 0
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_fused_exchange_collective_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


@pytest.mark.parametrize("world", [2, 3])
def test_edge_allgather_and_merge(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            spans = [vdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_single_process_passthrough():
    t = torch.tensor([5, 3, 9], dtype=torch.int64)
    assert vdist.allgather_varlen(t) is t
    assert vdist.merge_keys(t) is t  # one rank: the library's keys are already sorted, nothing to merge
    rp, ci = vdist.csr_from_keys(np.array([(0 << 32) | 4, (2 << 32) | 1, (2 << 32) | 7], dtype=np.uint64), 4)
    assert rp.tolist() == [0, 1, 1, 3, 3] and ci.tolist() == [4, 1, 7]
