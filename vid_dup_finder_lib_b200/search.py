"""`search` / `search_with_references` of vid_dup_finder_lib/src/video_hashing/video_dup_finder.rs, with the
comparison loops and the grouping rule running on the GPU behind the C ABI.  What stays here is what
include/vdf_b200.h leaves to the host language: the stable (duration, path) sort, the tolerance cast, the
index -> path mapping and MatchGroup construction."""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np

from . import _ffi
from .match_group import MatchGroup, MatchGroups
from .video_hash import HashTable, VideoHash, as_table


def search_indices(table: HashTable, tolerance: float, ctx: Optional[_ffi.Context] = None):
    """-> (group_ptr, member_idx): groups of indices into the table AS GIVEN (the library sorts internally)."""
    ctx = ctx or _ffi.default_context()
    blob, off = table.path_blob()
    return ctx.search(table.hashes, table.durations, blob, off, tolerance)


def search(hashes: Iterable[VideoHash] | HashTable, tolerance: float, ctx: Optional[_ffi.Context] = None) -> MatchGroups:
    """video_dup_finder.rs:7-13.  Every video is matched at most once; each group lists its matches in sorted
    order followed by the group's target (search_algorithm.rs:158-161), groups in the reference's order."""
    table = as_table(hashes)
    if len(table) == 0:  # search_algorithm.rs:88-90
        return MatchGroup.from_csr([], [0], [])
    gp, mm = search_indices(table, tolerance, ctx)
    return MatchGroup.from_csr(table.paths, gp, mm)  # MatchGroup::new(x).ok(), video_dup_finder.rs:11


def search_with_references(ref_hashes: Iterable[VideoHash] | HashTable, new_hashes: Iterable[VideoHash] | HashTable,
                           tolerance: float, ctx: Optional[_ffi.Context] = None) -> MatchGroups:
    """video_dup_finder.rs:19-46: one group per reference (in caller order) that matched at least one entry of
    new_hashes inside its duration slice; an entry may appear under many references (consume = false)."""
    ctx = ctx or _ffi.default_context()
    refs, cands = as_table(ref_hashes), as_table(new_hashes)
    if len(refs) == 0 or len(cands) == 0:
        return MatchGroup.from_csr([], [0], [])
    blob, off = cands.path_blob()
    rp, ci = ctx.search_with_references(refs.hashes, refs.durations, cands.hashes, cands.durations, blob, off, tolerance)
    return MatchGroup.from_csr(cands.paths, rp, ci, references=refs.paths)  # video_dup_finder.rs:38-43
