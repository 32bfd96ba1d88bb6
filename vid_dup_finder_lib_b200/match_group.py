"""`MatchGroup` of vid_dup_finder_lib/src/video_hashing/matches/match_group.rs."""
from __future__ import annotations

import itertools
from collections.abc import Sequence
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
    def from_csr(paths, ptr, idx, references=None) -> "MatchGroups":
        """Groups from a CSR of indices into `paths` (what the C ABI returns).  Row g is
        MatchGroup::new(paths[idx[ptr[g]:ptr[g+1]]]) -- or new_with_reference(references[g], ..) where rows that are
        empty are skipped (video_dup_finder.rs:11, :38-43)."""
        return MatchGroups(paths, ptr, idx, references)

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


class MatchGroups(Sequence):
    """The `Vec<MatchGroup>` a search returns, backed by the CSR the library hands over (group offsets + indices of
    the caller's entries): a MatchGroup object, with its path strings, comes into being when it is looked at.  A search
    over a million hashes returns ~10^5 groups; building them all eagerly cost a quarter of the call in round 1.
    Compares equal to any sequence of equal MatchGroups; `materialize()` gives the plain list."""

    __slots__ = ("_paths", "_ptr", "_idx", "_refs", "_rows")

    def __init__(self, paths, ptr, idx, references=None):
        import numpy as np

        self._paths, self._refs = paths, references
        self._ptr = np.asarray(ptr, dtype=np.int64)
        self._idx = np.asarray(idx, dtype=np.int64)
        sizes = np.diff(self._ptr)
        # MatchGroup::new needs two entries, new_with_reference one (match_group.rs:21-47): other rows do not exist
        self._rows = np.nonzero(sizes >= (2 if references is None else 1))[0]

    def __len__(self) -> int:
        return len(self._rows)

    def _make(self, r: int) -> MatchGroup:
        a, b = int(self._ptr[r]), int(self._ptr[r + 1])
        members = self._idx[a:b]
        if hasattr(self._paths, "dtype"):  # numpy array of paths
            dup = self._paths[members].tolist()
        else:
            dup = [self._paths[i] for i in members.tolist()]
        return MatchGroup(None if self._refs is None else self._refs[r], dup)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._make(int(r)) for r in self._rows[k]]
        return self._make(int(self._rows[k]))

    def __iter__(self) -> Iterator[MatchGroup]:
        for r in self._rows.tolist():
            yield self._make(r)

    def __eq__(self, other):
        try:
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    __hash__ = None

    def sizes(self):
        """entries per group (duplicates only, as MatchGroup.len()), without building any group"""
        import numpy as np

        return np.diff(self._ptr)[self._rows]

    def indices(self, k: int):
        """the caller's indices of group k's duplicates"""
        r = int(self._rows[k])
        return self._idx[int(self._ptr[r]):int(self._ptr[r + 1])]

    @property
    def csr(self):
        """(group_ptr, member_idx) exactly as the library returned them"""
        return self._ptr, self._idx

    def materialize(self) -> List[MatchGroup]:
        import gc

        was_enabled = gc.isenabled()
        gc.disable()  # ~10^5 small containers, no cycles: generational collections would triple the time
        try:
            return list(self)
        finally:
            if was_enabled:
                gc.enable()

    def __repr__(self):
        return f"MatchGroups({len(self)} groups)"
