"""`VideoHash` of vid_dup_finder_lib/src/video_hashing/video_hash.rs, and the struct-of-arrays table the GPU
search consumes."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from .definitions import HASH_BITS, HASH_WORDS, TOLERANCE_SCALING_FACTOR


@dataclass(frozen=True, order=True)
class VideoHash:
    """{hash: [usize; 16], src_path: PathBuf, duration: u32} (video_hash.rs:26-32); field order gives the same
    derived ordering."""

    hash: Tuple[int, ...]
    src_path: str = ""
    duration: int = 0

    def __post_init__(self):
        assert len(self.hash) == HASH_WORDS

    @staticmethod
    def from_words(words: Sequence[int], src_path: str = "", duration: int = 0) -> "VideoHash":
        return VideoHash(tuple(int(w) for w in words), os.fspath(src_path), int(duration))

    def hamming_distance(self, other: "VideoHash") -> int:
        """video_hash.rs:190-192,311-317: all 16 words, pad bits included.  A scalar accessor for callers that
        sort or inspect single pairs (search_output.rs:53); bulk comparison is the GPU search."""
        return sum((a ^ b).bit_count() for a, b in zip(self.hash, other.hash))

    def normalized_hamming_distance(self, other: "VideoHash") -> float:  # video_hash.rs:199-203
        return self.hamming_distance(other) / TOLERANCE_SCALING_FACTOR

    def raw_hash(self) -> List[bool]:  # video_hash.rs:207-216 : the 1000 meaningful bits, Lsb0
        return [bool((self.hash[b // 64] >> (b % 64)) & 1) for b in range(HASH_BITS)]

    hash_bits = raw_hash  # video_hash.rs:225-227

    @staticmethod
    def hash_frame_dimensions() -> Tuple[int, int]:  # video_hash.rs:219-222
        return (10, 10)

    def with_duration(self, duration: int) -> "VideoHash":  # test_util, video_hash.rs:252-256
        return VideoHash(self.hash, self.src_path, int(duration))

    def with_src_path(self, src_path) -> "VideoHash":  # video_hash.rs:258-263
        return VideoHash(self.hash, os.fspath(src_path), self.duration)

    def words(self) -> np.ndarray:
        return np.array(self.hash, dtype=np.uint64)


class HashTable:
    """Struct-of-arrays view of many VideoHashes: hashes [n,16] u64, durations [n] u32, paths list[str]."""

    __slots__ = ("hashes", "durations", "paths", "_blob")

    def __init__(self, hashes, durations, paths: Sequence[str]):
        self.hashes = np.ascontiguousarray(hashes, dtype=np.uint64).reshape(-1, 16)
        self.durations = np.ascontiguousarray(durations, dtype=np.uint32)
        self.paths = paths
        self._blob = None
        assert len(self.hashes) == len(self.durations) == len(paths)

    def path_blob(self):
        """(blob u8, offsets u64[n+1]): the src_paths as the C ABI takes them (include/vdf_b200.h, vdf_search); built
        once per table"""
        if self._blob is None:
            enc = [os.fsencode(p) for p in self.paths]
            off = np.zeros(len(enc) + 1, dtype=np.uint64)
            if enc:
                np.cumsum(np.fromiter(map(len, enc), dtype=np.uint64, count=len(enc)), out=off[1:])
            blob = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc else np.zeros(0, np.uint8)
            self._blob = (blob if len(blob) else np.zeros(1, np.uint8), off)
        return self._blob

    def __len__(self):
        return len(self.durations)

    @staticmethod
    def from_hashes(hashes: Iterable[VideoHash]) -> "HashTable":
        hs = hashes if isinstance(hashes, list) else list(hashes)
        if not hs:
            return HashTable(np.zeros((0, 16), np.uint64), np.zeros(0, np.uint32), [])
        return HashTable(np.array([h.hash for h in hs], dtype=np.uint64), np.array([h.duration for h in hs], dtype=np.uint32),
                         [h.src_path for h in hs])


def as_table(hashes) -> HashTable:
    return hashes if isinstance(hashes, HashTable) else HashTable.from_hashes(hashes)


# ---- Rust `Path` ordering (std::path::Path::cmp compares component lists) ------------------------------
def path_components(p: str):
    """Unix `Path::components()`: RootDir(1) < CurDir(2) < ParentDir(3) < Normal(4, bytes); '.' is dropped except
    as the leading component of a relative path; empty pieces are dropped."""
    b = os.fsencode(p)
    out = []
    if b.startswith(b"/"):
        out.append((1, b""))
        start = True
    else:
        start = False
    for k, piece in enumerate(b.split(b"/")):
        if piece == b"":
            continue
        if piece == b".":
            if k == 0 and not start:
                out.append((2, b""))
            continue
        out.append((3, b"") if piece == b".." else (4, piece))
    return out


def _is_simple(b: bytes) -> bool:
    # no empty / '.' / '..' components and no trailing slash: component order == byte order with '/' lowest
    if b.endswith(b"/"):  # a trailing separator is dropped, and "/" alone is the root component (Rust: "" < "/"), while numpy's
        return False      # bytes_ comparison ignores trailing NULs and would tie it with the empty path
    return not (b"//" in b or b"/./" in b or b"/../" in b or b.startswith((b"./", b"../")) or b.endswith((b"/.", b"/.."))
                or b in (b".", b".."))


def sort_order(durations: np.ndarray, paths: Sequence[str]) -> np.ndarray:
    """Permutation of Search::sort (search_algorithm.rs:55-61): stable, by (duration, src_path).  Pure-Python statement
    of the rule; the product path sorts natively (vdf_sort_order / vdf_search in csrc/host.cu) and
    tests/test_host_layer.py holds the two against each other."""
    n = len(paths)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    bs = [os.fsencode(p) for p in paths]
    if all(_is_simple(b) for b in bs):
        # mapping '/' to NUL makes plain byte order equal component order (NUL never occurs in a Unix path)
        keys = np.array([b.replace(b"/", b"\x00") for b in bs], dtype=np.bytes_)
        return np.lexsort((keys, np.asarray(durations, dtype=np.uint32)))
    d = [int(v) for v in durations]
    comps = [path_components(p) for p in paths]
    return np.array(sorted(range(n), key=lambda i: (d[i], comps[i])), dtype=np.int64)
