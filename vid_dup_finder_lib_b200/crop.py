"""`Crop` of vid_dup_finder_common/src/crop.rs (the parts the hashing path uses, plus the tested helpers)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, Tuple


@dataclass(frozen=True, order=True)
class Crop:
    orig_res: Tuple[int, int]
    left: int
    right: int
    top: int
    bottom: int

    @staticmethod
    def from_edge_offsets(orig_res, left, right, top, bottom) -> "Crop":  # crop.rs:14-30
        assert (left + right) < orig_res[0]
        assert (top + bottom) < orig_res[1]
        return Crop(tuple(orig_res), left, right, top, bottom)

    @staticmethod
    def from_topleft_and_dims(orig_res, x, y, width, height) -> "Crop":  # crop.rs:32-50
        ow, oh = orig_res
        return Crop((ow, oh), x, ow - width - x, y, oh - height - y)

    def union(self, other: "Crop") -> "Crop":  # crop.rs:53-68 : per-side minimum
        return Crop.from_edge_offsets(self.orig_res, min(self.left, other.left), min(self.right, other.right),
                                      min(self.top, other.top), min(self.bottom, other.bottom))

    def as_view_args(self) -> Tuple[int, int, int, int]:  # crop.rs:92-103
        ow, oh = self.orig_res
        w, h = ow - (self.left + self.right), oh - (self.top + self.bottom)
        assert w >= 0 and h >= 0
        return (self.left, self.top, w, h)

    def width(self) -> int:
        return self.orig_res[0] - (self.left + self.right)

    def height(self) -> int:
        return self.orig_res[1] - (self.top + self.bottom)

    def area(self) -> int:
        return self.width() * self.height()

    def is_uncropped(self) -> bool:
        return self.left == 0 and self.right == 0 and self.top == 0 and self.bottom == 0

    def enumerate_coords(self) -> Iterator[Tuple[int, int]]:  # crop.rs:121-133 (x outer, y inner)
        ow, oh = self.orig_res
        for x in range(self.left, ow - self.right):
            for y in range(self.top, oh - self.bottom):
                yield (x, y)

    def enumerate_coords_excluded(self) -> Iterator[Tuple[int, int]]:  # crop.rs:135-160
        ow, oh = self.orig_res
        xs = [(0, self.left), (self.left, ow - self.right), (ow - self.right, ow)]
        ys = [(0, self.top), (self.top, oh - self.bottom), (oh - self.bottom, oh)]
        # clockwise from top-left, the middle cell (the kept window) is skipped
        for xi, yi in [(0, 0), (1, 0), (2, 0), (2, 1), (0, 2), (1, 2), (2, 2), (0, 1)]:
            for x in range(*xs[xi]):
                for y in range(*ys[yi]):
                    yield (x, y)
