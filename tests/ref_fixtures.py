"""Python re-expression of the reference's test fixtures.

`VideoHash::{random_hash, hash_with_spatial_distance, full_hash, empty_hash}` follow
vid_dup_finder_lib/src/video_hashing/video_hash.rs:252-306; `HashesWithDistance` and
`HashesWithDistanceSet` follow vid_dup_finder_lib/tests/test_find_all.rs:14-132.  The reference
asserts only group counts/sizes on these fixtures, so any PRNG reproduces them.
Hashes are numpy uint64[16] rows.
"""
from __future__ import annotations

import itertools

import numpy as np

WORDS = 16
HASH_BITS = 1000
MASK40 = (1 << 40) - 1


def popcount64(a: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(a).view(np.uint8), axis=-1).sum(axis=-1)


def hamming(a: np.ndarray, b: np.ndarray) -> int:
    return int(popcount64(a ^ b))


def random_hash(rng: np.random.Generator) -> np.ndarray:
    """video_hash.rs:293-306: 1000 random bits, pad bits 1000..1023 zero."""
    h = rng.integers(0, 1 << 64, WORDS, dtype=np.uint64)
    h[15] &= np.uint64(MASK40)
    return h


def full_hash() -> np.ndarray:
    return np.full(WORDS, np.uint64(0xFFFFFFFFFFFFFFFF))  # video_hash.rs:265-267 (pad bits set too)


def empty_hash() -> np.ndarray:
    return np.zeros(WORDS, dtype=np.uint64)


def hash_with_spatial_distance(h: np.ndarray, target: int, rng: np.random.Generator) -> np.ndarray:
    """video_hash.rs:272-291: a hash at exactly `target` bits from `h`, flipped bits drawn from all 16x64
    positions (pad bits included).  The reference random-walks single bit flips until the distance is
    reached (minutes of Python for the 600-bit separations of test_find_with_refs); flipping `target`
    distinct random positions meets the same postcondition (:289) directly."""
    ret = h.copy()
    for pos in rng.choice(WORDS * 64, size=target, replace=False):
        ret[int(pos) // 64] ^= np.uint64(1 << (int(pos) % 64))
    assert hamming(h, ret) == target
    return ret


class HashesWithDistance:
    """test_find_all.rs:14-66"""

    def __init__(self, start_hash, distance_from_start, num_hashes, rng):
        self.start_hash = start_hash
        self.members_ = [hash_with_spatial_distance(start_hash, distance_from_start, rng) for _ in range(num_hashes)]
        m = np.stack(self.members_) if self.members_ else np.zeros((0, WORDS), np.uint64)
        if len(m) > 1:  # sanity check :44-50 (triangle inequality)
            d = popcount64(m[:, None, :] ^ m[None, :, :])
            assert d.max() <= 2 * distance_from_start

    def members(self, rng):
        ret = list(self.members_)
        rng.shuffle(ret)
        return ret


class HashesWithDistanceSet:
    """test_find_all.rs:68-132"""

    def __init__(self, num_groups, hashes_per_group, intergroup_distance, intragroup_distance, rng):
        assert intragroup_distance * 2 < intergroup_distance
        assert (19 * 64) // num_groups > intergroup_distance
        start = random_hash(rng)
        cur = 0
        self.groups = []
        for _ in range(num_groups):
            gstart = hash_with_spatial_distance(start, cur, rng)
            cur += intergroup_distance
            self.groups.append(HashesWithDistance(gstart, intragroup_distance, hashes_per_group, rng))
            hashes_per_group += 10

    def all_members(self, rng):
        allm = list(itertools.chain.from_iterable(g.members(rng) for g in self.groups))
        rng.shuffle(allm)
        return allm
