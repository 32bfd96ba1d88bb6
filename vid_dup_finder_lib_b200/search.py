"""`search` / `search_with_references` of vid_dup_finder_lib/src/video_hashing/video_dup_finder.rs, with the
comparison loops and the grouping rule running on the GPU behind the C ABI.  What stays here is what
include/vdf_b200.h leaves to the host language: the stable (duration, path) sort, the tolerance cast, the
index -> path mapping and MatchGroup construction."""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np

from . import _ffi
from .definitions import tolerance_to_int
from .match_group import MatchGroup
from .video_hash import HashTable, VideoHash, as_table, sort_order


def search_indices(table: HashTable, tolerance: float, ctx: Optional[_ffi.Context] = None):
    """-> (order, group_ptr, member_idx): groups of indices into the SORTED table; order maps back to the input."""
    ctx = ctx or _ffi.default_context()
    order = sort_order(table.durations, table.paths)
    gp, mm = ctx.search_self_groups(table.hashes[order], table.durations[order], tolerance_to_int(tolerance))
    return order, gp, mm


def search(hashes: Iterable[VideoHash] | HashTable, tolerance: float, ctx: Optional[_ffi.Context] = None) -> List[MatchGroup]:
    """video_dup_finder.rs:7-13.  Every video is matched at most once; each group lists its matches in sorted
    order followed by the group's target (search_algorithm.rs:158-161), groups in the reference's order."""
    table = as_table(hashes)
    if len(table) == 0:  # search_algorithm.rs:88-90
        return []
    order, gp, mm = search_indices(table, tolerance, ctx)
    paths = table.paths
    out = []
    for g in range(len(gp) - 1):
        members = [paths[order[k]] for k in mm[gp[g]:gp[g + 1]]]
        if len(members) >= 2:  # MatchGroup::new(x).ok(), video_dup_finder.rs:11
            out.append(MatchGroup.new(members))
    return out


def search_with_references(ref_hashes: Iterable[VideoHash] | HashTable, new_hashes: Iterable[VideoHash] | HashTable,
                           tolerance: float, ctx: Optional[_ffi.Context] = None) -> List[MatchGroup]:
    """video_dup_finder.rs:19-46: one group per reference (in caller order) that matched at least one entry of
    new_hashes inside its duration slice; an entry may appear under many references (consume = false)."""
    ctx = ctx or _ffi.default_context()
    refs, cands = as_table(ref_hashes), as_table(new_hashes)
    if len(refs) == 0 or len(cands) == 0:
        return []
    order = sort_order(cands.durations, cands.paths)
    rp, ci = ctx.search_refs(cands.hashes[order], cands.durations[order], refs.hashes, refs.durations,
                             tolerance_to_int(tolerance))
    out = []
    for r in range(len(refs)):
        if rp[r + 1] > rp[r]:  # video_dup_finder.rs:38-43
            out.append(MatchGroup.new_with_reference(refs.paths[r], [cands.paths[order[k]] for k in ci[rp[r]:rp[r + 1]]]))
    return out
