#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/dist_check.py
Every rank runs the sharded `search`, `search_with_references` and stack hashing over NCCL and compares the merged
result with the CPU oracle (test infrastructure) on the same inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vid_dup_finder_lib_b200 as vdf  # noqa: E402
from oracle import vdf_oracle as o  # noqa: E402
from tests import synth  # noqa: E402
from vid_dup_finder_lib_b200 import _ffi  # noqa: E402
from vid_dup_finder_lib_b200 import dist as vdist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = _ffi.default_context()
    ok = True

    n = 30000
    H, _ = synth.planted_hashes(n, seed=11, dup_frac_den=4)
    dur = synth.lognormal_durations(n, seed=11)
    paths = synth.paths(n)
    table = vdf.HashTable(H, dur, paths)

    def check_searches(label):
        ok = True
        # search: planted duplicates, log-normal durations
        for tol in (0.1, 0.35):
            groups = vdist.search(table, tol, ctx=ctx)
            order = o.sort_order(dur, paths)
            gp, mm = o.search_self(H[order], dur[order], o.tolerance_int(tol))
            want = [[paths[order[k]] for k in mm[gp[g]:gp[g + 1]]] for g in range(len(gp) - 1)]
            got = [list(g.duplicates()) for g in groups]
            ok &= got == want
            if rank == 0:
                print(f"[{label}] search tol={tol}: {len(got)} groups, parity={'ok' if got == want else 'MISMATCH'}")

        # search_with_references: candidate slices per rank
        refs = vdf.HashTable(H[::97][:200], dur[::97][:200], ["ref/%04d" % i for i in range(200)])
        got = vdist.search_with_references(refs, table, 0.3, ctx=ctx)
        order = o.sort_order(dur, paths)
        rp, ci = o.search_refs(H[order], dur[order], refs.hashes, refs.durations, o.tolerance_int(0.3))
        want = [(refs.paths[r], [paths[order[k]] for k in ci[rp[r]:rp[r + 1]]]) for r in range(200) if rp[r + 1] > rp[r]]
        got2 = [(g.reference(), list(g.duplicates())) for g in got]
        ok &= got2 == want
        if rank == 0:
            print(f"[{label}] search_with_references: {len(got2)} groups, parity={'ok' if got2 == want else 'MISMATCH'}")
        return ok

    ok &= check_searches("nccl all-gather")
    # the same searches with the exchange fused into the pair kernel (appends over NVLink peer memory); the first capacity
    # is too small on purpose: every rank sees the overflow with the same count and re-allocates with the others
    vdist.enable_peer_exchange(ctx, capacity=100)
    ok &= check_searches("peer-memory exchange")
    ok &= ctx.peer_capacity > 100
    for _ in range(20):  # alternate the two halves of the exchange buffer a few more times
        got = vdist.search(table, 0.2, ctx=ctx)
    want_n = len(got)
    ok &= all(len(vdist.search(table, 0.2, ctx=ctx)) == want_n for _ in range(3))
    vdist.disable_peer_exchange(ctx)

    # hashing: stacks sharded by rank, hashes all-gathered in rank order
    per = 6
    st = synth.frame_stacks(per, 640, 360, seed=5, device="cuda", first_id=rank * per)
    allh, status = vdist.hash_stacks_sharded(ctx, st, _ffi.make_descs(per, 640, 360), _ffi.CROPDETECT_LETTERBOX)
    mine = allh[rank * per:(rank + 1) * per].cpu().numpy().view(np.uint64)
    want_h = np.stack([o.hash_stack(st[s].cpu().numpy(), 1)[1] for s in range(per)])
    ok &= allh.shape[0] == per * world and np.array_equal(mine, want_h) and not status.any()
    if rank == 0:
        print(f"hash shard: {allh.shape[0]} hashes gathered, parity={'ok' if np.array_equal(mine, want_h) else 'MISMATCH'}")

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK", "PASSED" if flag.item() == 1 else "FAILED", f"(world={world})")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
