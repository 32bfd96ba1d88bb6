"""`VideoHashBuilder` / `CreationOptions` / `Error` of vid_dup_finder_lib (video_hash_builder.rs, video_hashing/mod.rs)
for the part of hashing that runs on the GPU: everything after the frames are decoded.  Decoding (ffmpeg /
gstreamer, video_hash_builder.rs:85-167) stays with the caller, who hands over the gray u8 frames."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _ffi
from .definitions import (DCT_SIZE, DEFAULT_VID_HASH_DURATION, DEFAULT_VID_HASH_SKIP_FORWARD, Cropdetect)
from .video_hash import VideoHash


class Error(Exception):
    """vid_dup_finder_lib::Error (video_hashing/mod.rs:17-28)"""


class NotVideo(Error):
    def __str__(self):
        return "File is not a video"


class VidProc(Error):
    def __str__(self):
        return f"Video processing error: {self.args[0] if self.args else ''}"


class NotEnoughFrames(Error):
    def __str__(self):
        return "Could not extract enough frames"


@dataclass(frozen=True)
class CreationOptions:
    """video_hash_builder.rs:17-63; skip_forward_amount and duration steer the decoder's frame schedule (host)."""

    skip_forward_amount: float = DEFAULT_VID_HASH_SKIP_FORWARD
    duration: float = DEFAULT_VID_HASH_DURATION
    cropdetect: Cropdetect = Cropdetect.LETTERBOX


def frame_schedule(vid_duration: float, opts: CreationOptions = CreationOptions()):
    """build_frame_reader's fps/seek policy (video_hash_builder.rs:104-146) -> ((fps_num, fps_den), seek_seconds)."""
    if vid_duration < 2.0:
        fps, seek = 64.0, 0.0
    elif vid_duration < opts.duration:
        fps, seek = 64.0 / (vid_duration - 2.0), 0.0
    elif vid_duration < opts.skip_forward_amount + opts.duration:
        fps, seek = 64.0 / opts.duration, vid_duration - opts.duration - 2.0
    else:
        fps, seek = 64.0 / opts.duration, opts.skip_forward_amount
    return (int(fps * 16384.0), 16384), seek


def _cropdetect_code(c: Cropdetect) -> int:
    return {Cropdetect.NONE: _ffi.CROPDETECT_NONE, Cropdetect.LETTERBOX: _ffi.CROPDETECT_LETTERBOX,
            Cropdetect.MOTION: _ffi.CROPDETECT_MOTION}[c]  # Motion: csrc/motion.cu


class VideoHashBuilder:
    """video_hash_builder.rs:70-83.  `hash_frames` is gen_hash's compute tail for one video; `hash_many` batches
    stacks so that the GPU sees thousands of frames per launch."""

    def __init__(self, options: CreationOptions = CreationOptions(), ctx: Optional[_ffi.Context] = None):
        self.options = options
        self._ctx = ctx

    @staticmethod
    def from_options(options: CreationOptions) -> "VideoHashBuilder":
        return VideoHashBuilder(options)

    @property
    def ctx(self) -> _ffi.Context:
        if self._ctx is None:
            self._ctx = _ffi.default_context()
        return self._ctx

    def hash_frames(self, frames: Sequence[np.ndarray], src_path: str, duration_secs: int) -> VideoHash:
        res = self.hash_many([frames], [src_path], [duration_secs])[0]
        if isinstance(res, Error):
            raise res
        return res

    def hash_many(self, stacks: Sequence[Sequence[np.ndarray]], src_paths: Sequence[str],
                  durations: Sequence[int]) -> List[Union[VideoHash, Error]]:
        """stacks[i] = the gray frames of video i (any number; the first 16 are used, video_hash_builder.rs:164).
        Returns, per video, a VideoHash or the Error the reference would return."""
        n = len(stacks)
        descs = np.zeros(n, dtype=_ffi.STACK_DESC_DTYPE)
        chunks, off = [], 0
        for i, st in enumerate(stacks):
            fr = [np.ascontiguousarray(f, dtype=np.uint8) for f in list(st)[:DCT_SIZE]]
            d = descs[i]
            d["n_frames"] = len(fr)
            if not fr:
                continue
            h, w = fr[0].shape
            if any(f.shape != (h, w) for f in fr):  # are_all_frames_same_size, video_hash_builder.rs:169-186
                d["flags"] = _ffi.STACK_FLAG_MIXED_SIZES
                continue
            d["offset"], d["frame_stride"], d["width"], d["height"], d["pitch"] = off, w * h, w, h, w
            chunks.extend(fr)
            off += w * h * len(fr)
        buf = np.concatenate([c.reshape(-1) for c in chunks]) if chunks else np.zeros(1, np.uint8)
        hashes, status, _ = self.ctx.hash_stacks(buf, descs, _cropdetect_code(self.options.cropdetect))
        out: List[Union[VideoHash, Error]] = []
        for i in range(n):
            if status[i] == _ffi.STACK_OK:
                out.append(VideoHash.from_words(hashes[i], src_paths[i], durations[i]))
            elif status[i] == _ffi.STACK_VIDPROC:
                out.append(VidProc("frames not all same size"))
            else:
                out.append(NotEnoughFrames())
        return out
