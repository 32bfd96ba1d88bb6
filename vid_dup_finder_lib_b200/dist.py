"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).

How the two paths shard:
  * hashing            - by stack, contiguous ranges, no exchange while hashing; one all-gather of the 128-byte
                         hashes afterwards so that every rank holds the table the search needs.
  * search_self        - every rank holds the full sorted table; the (row tile, chunk) units of the pair matrix are
                         dealt block-cyclically to the ranks inside the kernel (vdf_ctx_set_shard); ONE exchange of
                         edges; the greedy grouping is replicated.  The exchange is fused into the pair kernel when
                         `enable_peer_exchange` has been called: every match is appended to every rank's buffer over
                         NVLink peer memory by the kernel that found it (include/vdf_b200.h: vdf_peer_*).  Without
                         it: a variable-length NCCL all-gather of the per-rank edge lists.
  * search_with_refs   - the sorted candidate table is cut into contiguous slices, the references are replicated;
                         per-rank (ref, cand) keys are all-gathered and merged by a sort.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi
from .definitions import tolerance_to_int
from .match_group import MatchGroup, MatchGroups
from .video_hash import HashTable, as_table


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) of n items for `rank`; sizes differ by at most one"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allgather_varlen(local: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenation over ranks (in rank order) of 1-D tensors of different lengths."""
    rank, world = world_info(group)
    if world == 1:
        return local
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    if m == 0:
        return local[:0]
    padded = torch.zeros(m, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def merge_keys(local_keys: torch.Tensor, group=None) -> torch.Tensor:
    """all ranks' u64 keys (carried as int64; indices < 2^31 keep them non-negative), globally sorted; each rank's
    own keys arrive sorted from the library, so a single rank has nothing to merge"""
    if world_info(group)[1] == 1:
        return local_keys  # the library already returns its keys sorted
    allk = allgather_varlen(local_keys, group)
    return torch.sort(allk).values if allk.numel() else allk


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()


def _run_growing(fn, device, initial: int = 1 << 22, ctx=None):
    cap = initial
    while True:
        if ctx is not None:
            keys = _key_buffer(ctx, cap, device)
            cap = keys.numel()
        else:
            keys = torch.empty(cap, dtype=torch.int64, device=device)
            _sync(device)
        cnt = fn(keys.data_ptr(), cap)
        if cnt >= 0:
            return keys[:cnt]
        cap = -cnt + 1024


def enable_peer_exchange(ctx: _ffi.Context, group=None, capacity: int = 1 << 22) -> bool:
    """Set up the fused edge exchange (one buffer per rank, mapped by all peers through CUDA IPC).  Collective: every rank
    of the group calls it with the same capacity (matches of ALL ranks that must fit).  False on a single rank."""
    rank, world = world_info(group)
    if world < 2:
        return False
    if world > 8:
        raise ValueError("the peer exchange covers the <= 8 GPUs of one node")
    handle = ctx.peer_alloc(capacity)
    handles: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    ctx.peer_open(rank, world, b"".join(handles))
    dist.barrier(group=group)  # every rank has mapped every buffer before anyone appends
    return True


def disable_peer_exchange(ctx: _ffi.Context, group=None):
    if ctx.peer_world:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier(group=group)  # nobody is still appending to a buffer about to be freed
        ctx.peer_close()


def _key_buffer(ctx: _ffi.Context, cap: int, device) -> torch.Tensor:
    """the context's key buffer (the result of a search is a view of it, valid until the context's next search): allocating
    4 M keys and synchronising for every search cost a fifth of a millisecond per step"""
    buf = getattr(ctx, "_key_buf", None)
    if buf is None or buf.numel() < cap or buf.device != torch.device(device):
        buf = torch.empty(cap, dtype=torch.int64, device=device)
        _sync(device)
        ctx._key_buf = buf
    return buf


def _run_exchange(ctx: _ffi.Context, fn, device, group):
    """fn with the library's "exchange" option on: returns the sorted keys of ALL ranks.  An overflow is reported with the same
    global count on every rank, so all ranks re-allocate (collectively) and retry together."""
    while True:
        cap = ctx.peer_capacity
        keys = _key_buffer(ctx, cap, device)
        ctx.set_option("exchange", 1)
        try:
            cnt = fn(keys.data_ptr(), cap)
        finally:
            ctx.set_option("exchange", 0)
        if cnt >= 0:
            return keys[:cnt]
        need = -cnt
        disable_peer_exchange(ctx, group)
        enable_peer_exchange(ctx, group, need + need // 16 + 1024)


def search_self_keys(ctx: _ffi.Context, d_hash, d_dur, tol_int: int, group=None, n: Optional[int] = None,
                     device=None) -> torch.Tensor:
    """This rank's share of the pair matrix and the edge exchange: sorted (i << 32 | j) keys of ALL ranks.
    d_hash / d_dur: the sorted table resident in HBM, as torch tensors or as raw device pointers (then pass n, device);
    or d_hash = a prepared _ffi.Table (`Search::from` done once: nothing is re-packed per call), d_dur ignored, pass device."""
    rank, world = world_info(group)
    table = d_hash if isinstance(d_hash, _ffi.Table) else None
    if table is not None:
        pass
    elif isinstance(d_hash, torch.Tensor):
        n, device, p_hash, p_dur = d_dur.numel(), d_hash.device, d_hash.data_ptr(), d_dur.data_ptr()
    else:
        p_hash, p_dur = int(d_hash), int(d_dur)
    fused = world > 1 and ctx.peer_world == world
    ctx.set_shard(rank, world)
    try:
        if table is not None:
            run = lambda p, cap: table.search_self_device(tol_int, p, cap)  # noqa: E731
        else:
            run = lambda p, cap: ctx.search_self_device(p_hash, p_dur, n, tol_int, p, cap)  # noqa: E731
        if fused:
            return _run_exchange(ctx, run, device, group)
        local = _run_growing(run, device, ctx=ctx)
    finally:
        ctx.set_shard(0, 1)
    return merge_keys(local, group)


def search_refs_keys(ctx: _ffi.Context, d_cand_slice, d_cand_dur_slice, cand_base: int,
                     d_refs: torch.Tensor, d_ref_dur: torch.Tensor, tol_int: int, group=None, allow_fused: bool = True) -> torch.Tensor:
    """This rank's slice of the sorted candidate table against all references: sorted (ref << 32 | cand) keys of ALL ranks.
    d_cand_slice: a tensor (with d_cand_dur_slice) or a prepared _ffi.Table of the slice.  With the exchange on, every rank
    reaches its closing barrier even when its slice is empty."""
    if isinstance(d_cand_slice, _ffi.Table):
        run = lambda p, cap: d_cand_slice.search_refs_device(cand_base, d_refs.data_ptr(), d_ref_dur.data_ptr(),  # noqa: E731
                                                             d_ref_dur.numel(), tol_int, p, cap)
    else:
        run = lambda p, cap: ctx.search_refs_device(d_cand_slice.data_ptr(), d_cand_dur_slice.data_ptr(),  # noqa: E731
                                                    d_cand_dur_slice.numel(), cand_base, d_refs.data_ptr(), d_ref_dur.data_ptr(),
                                                    d_ref_dur.numel(), tol_int, p, cap)
    world = world_info(group)[1]
    if allow_fused and world > 1 and ctx.peer_world == world:
        return _run_exchange(ctx, run, d_refs.device, group)
    return merge_keys(_run_growing(run, d_refs.device, ctx=ctx), group)


def csr_from_keys(keys: np.ndarray, n_rows: int) -> Tuple[np.ndarray, np.ndarray]:
    """sorted (row << 32 | col) keys -> (row_ptr, col_idx)"""
    keys = np.asarray(keys).astype(np.uint64)
    rows = (keys >> np.uint64(32)).astype(np.int64)
    rp = np.zeros(n_rows + 1, dtype=np.uint64)
    np.add.at(rp, rows + 1, 1)
    return np.cumsum(rp).astype(np.uint64), keys & np.uint64(0xFFFFFFFF)


def _to_dev(a: np.ndarray, device) -> torch.Tensor:
    t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32) if a.dtype == np.uint32 else a)
    return t.to(device, non_blocking=False)


def search(hashes, tolerance: float, ctx: Optional[_ffi.Context] = None, group=None) -> MatchGroups:
    """`search` (video_dup_finder.rs:7-13) over all ranks of the process group; every rank gets the full result."""
    ctx = ctx or _ffi.default_context()
    table = as_table(hashes)
    n = len(table)
    if n == 0:
        return MatchGroup.from_csr([], [0], [])
    if world_info(group)[1] == 1:  # one process: the whole function is one C-ABI call (vdf_search; also a multi-device context)
        from .search import search as search_one

        return search_one(table, tolerance, ctx)
    dev = torch.device("cuda", ctx.device)
    # Rank 0 does the host side of vdf_search once (sort keys cut by host threads, radix sort on its GPU, the hashes through
    # pinned memory) and broadcasts the sorted table and the permutation over NVLink; all ranks then search their shares.
    # Nothing else is serial: groups come back as a CSR of the caller's indices and become objects when looked at.
    rank = world_info(group)[0]
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)):
        d_hash = torch.empty((n, 16), dtype=torch.int64, device=dev)
        d_dur = torch.empty(n, dtype=torch.int32, device=dev)
        d_order = torch.empty(n, dtype=torch.int32, device=dev)
        if rank == 0:
            order, _, _ = ctx.stage_sorted(table.hashes, table.durations, *table.path_blob(), d_hash_dst=d_hash.data_ptr(),
                                           d_dur_dst=d_dur.data_ptr())
            d_order.copy_(torch.from_numpy(order.astype(np.int32)), non_blocking=False)
        dist.broadcast(d_hash, 0, group=group)
        dist.broadcast(d_dur, 0, group=group)
        dist.broadcast(d_order, 0, group=group)
        keys = search_self_keys(ctx, d_hash, d_dur, tolerance_to_int(tolerance), group)
        torch.cuda.current_stream().synchronize()
        # sorted positions -> the caller's indices on the device, as the CSR is written
        gp, mm = ctx.group_greedy_device(n, keys.data_ptr(), keys.numel(), d_remap=d_order.data_ptr())
    return MatchGroup.from_csr(table.paths, gp, mm)


def search_with_references(ref_hashes, new_hashes, tolerance: float, ctx: Optional[_ffi.Context] = None,
                           group=None) -> MatchGroups:
    """`search_with_references` (video_dup_finder.rs:19-46) with the sorted candidate table sliced over the ranks."""
    ctx = ctx or _ffi.default_context()
    refs, cands = as_table(ref_hashes), as_table(new_hashes)
    if len(refs) == 0 or len(cands) == 0:
        return MatchGroup.from_csr([], [0], [])
    rank, world = world_info(group)
    if world == 1:
        from .search import search_with_references as one

        return one(refs, cands, tolerance, ctx)
    dev = torch.device("cuda", ctx.device)
    n = len(cands)
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)):
        d_hash = torch.empty((n, 16), dtype=torch.int64, device=dev)
        d_dur = torch.empty(n, dtype=torch.int32, device=dev)
        d_order = torch.empty(n, dtype=torch.int64, device=dev)
        if rank == 0:  # Search::from(new_hashes) once, on rank 0's GPU; the sorted table travels over NVLink
            order, _, _ = ctx.stage_sorted(cands.hashes, cands.durations, *cands.path_blob(), d_hash_dst=d_hash.data_ptr(),
                                           d_dur_dst=d_dur.data_ptr())
            d_order.copy_(torch.from_numpy(order), non_blocking=False)
        dist.broadcast(d_hash, 0, group=group)
        dist.broadcast(d_dur, 0, group=group)
        dist.broadcast(d_order, 0, group=group)
        b, e = shard_range(n, rank, world)
        d_r = _to_dev(refs.hashes, dev)
        d_rd = _to_dev(refs.durations, dev)
        keys = search_refs_keys(ctx, d_hash[b:e], d_dur[b:e], b, d_r, d_rd, tolerance_to_int(tolerance), group)
        order = d_order.cpu().numpy()
    rp, ci = csr_from_keys(keys.cpu().numpy(), len(refs))
    return MatchGroup.from_csr(cands.paths, rp, order[ci.astype(np.int64)], references=refs.paths)


def hash_stacks_sharded(ctx: _ffi.Context, d_frames: torch.Tensor, descs: np.ndarray, cropdetect: int, group=None):
    """Hash this rank's stacks (descs index into d_frames, which holds only local stacks), then all-gather the
    128-byte hashes so that every rank holds the table in rank order.  -> ([n_total,16] int64 tensor, status)"""
    n = len(descs)
    out = torch.zeros((n, 16), dtype=torch.int64, device=d_frames.device)
    torch.cuda.current_stream().synchronize()
    status, _ = ctx.hash_stacks_device(d_frames.data_ptr(), descs, cropdetect, out.data_ptr())
    allh = allgather_varlen(out.reshape(-1), group).reshape(-1, 16)
    return allh, status
