"""`MatchGroup` of vid_dup_finder_lib/src/video_hashing/matches/match_group.rs."""
from __future__ import annotations

import itertools
from typing import Iterable, Iterator, List, Optional


class TooFewEntries(ValueError):
    pass


class MatchGroup:
    """A group of duplicate videos found by `search` or `search_with_references` (match_group.rs:9-13)."""

    __slots__ = ("_reference", "_duplicates")

    def __init__(self, reference: Optional[str], duplicates: List[str]):
        self._reference = reference
        self._duplicates = duplicates

    @staticmethod
    def new(entries: Iterable[str]) -> "MatchGroup":  # match_group.rs:21-30
        d = list(entries)
        if len(d) < 2:
            raise TooFewEntries()
        return MatchGroup(None, d)

    @staticmethod
    def new_with_reference(reference: str, entries: Iterable[str]) -> "MatchGroup":  # match_group.rs:35-47
        d = list(entries)
        if not d:
            raise TooFewEntries()
        return MatchGroup(reference, d)

    @staticmethod
    def from_csr(paths, ptr, idx, references=None) -> List["MatchGroup"]:
        """Groups from a CSR of indices into `paths` (what the C ABI returns).  Row g becomes
        MatchGroup::new(paths[idx[ptr[g]:ptr[g+1]]]) -- or new_with_reference(references[g], ..) where rows that are
        empty are skipped (video_dup_finder.rs:11, :38-43).  One bulk gather instead of a Python loop per path."""
        import numpy as np

        ptr = np.asarray(ptr, dtype=np.int64)
        idx = np.asarray(idx, dtype=np.int64)
        if isinstance(paths, np.ndarray):
            flat = paths[idx].tolist()
        else:
            flat = list(map(paths.__getitem__, idx.tolist()))
        sizes = np.diff(ptr)
        rows = np.nonzero(sizes >= (2 if references is None else 1))[0]
        a, b = ptr[rows].tolist(), ptr[rows + 1].tolist()
        import gc

        was_enabled = gc.isenabled()
        gc.disable()  # ~10^5 small containers, no cycles: generational collections would triple the time
        try:
            if references is None:
                return [MatchGroup(None, flat[x:y]) for x, y in zip(a, b)]
            return [MatchGroup(references[r], flat[x:y]) for r, x, y in zip(rows.tolist(), a, b)]
        finally:
            if was_enabled:
                gc.enable()

    def __len__(self) -> int:  # match_group.rs:51-53
        return len(self._duplicates)

    def len(self) -> int:
        return len(self._duplicates)

    def reference(self) -> Optional[str]:  # match_group.rs:57-59
        return self._reference

    def duplicates(self) -> Iterator[str]:  # match_group.rs:62-64
        return iter(self._duplicates)

    def contained_paths(self) -> Iterator[str]:  # match_group.rs:68-82 : duplicates, then the reference
        yield from self._duplicates
        if self._reference is not None:
            yield self._reference

    def dup_combinations(self) -> List["MatchGroup"]:  # match_group.rs:89-105
        if self._reference is not None:
            return [MatchGroup.new_with_reference(self._reference, [d]) for d in self._duplicates]
        return [MatchGroup.new([a, b]) for a, b in itertools.combinations(self._duplicates, 2)]

    def _key(self):
        return (self._reference is not None, self._reference or "", self._duplicates)

    def __eq__(self, other):
        return isinstance(other, MatchGroup) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __hash__(self):
        return hash((self._reference, tuple(self._duplicates)))

    def __repr__(self):
        return f"MatchGroup(reference={self._reference!r}, duplicates={self._duplicates!r})"
