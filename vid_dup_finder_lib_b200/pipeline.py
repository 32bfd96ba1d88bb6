"""Batch-oriented hashing: the app's per-file loop
(vid_dup_finder_app/src/video_hash_filesystem_cache/video_hash_filesystem_cache.rs:237-257: each rayon worker decodes AND
hashes one file) turned into decode threads -> pinned batch buffers -> GPU batches -> results, behind the C ABI
(vdf_pipeline_*, csrc/pipeline.cu)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _ffi
from .definitions import DCT_SIZE
from .video_hash_builder import CreationOptions, Error, NotEnoughFrames, VidProc, _cropdetect_code


class _Result(C.Structure):
    _fields_ = [("tag", C.c_uint64), ("status", C.c_int32), ("crop", C.c_uint32 * 4), ("hash", C.c_uint64 * 16)]


def _lib():
    L = _ffi.lib()
    vp = C.c_void_p
    L.vdf_pipeline_create.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.vdf_pipeline_push.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.vdf_pipeline_flush.argtypes = [vp]
    L.vdf_pipeline_poll.argtypes = [vp, C.POINTER(_Result), C.c_uint32, C.POINTER(C.c_uint32), C.c_int]
    L.vdf_pipeline_error.argtypes = [vp]
    L.vdf_pipeline_error.restype = C.c_char_p
    L.vdf_pipeline_destroy.argtypes = [vp]
    L.vdf_pipeline_destroy.restype = None
    return L


class HashPipeline:
    """push(tag, frames) from any number of decode threads (the call copies the frames into pinned memory and releases the
    GIL while it does); results() from a collector.  One worker thread inside the library owns the context while the
    pipeline lives.  `with HashPipeline(...) as p:` flushes and closes."""

    def __init__(self, options: CreationOptions = CreationOptions(), ctx: Optional[_ffi.Context] = None,
                 max_batch_stacks: int = 64, batch_bytes: int = 1 << 30):
        self._ctx = ctx or _ffi.default_context()
        self._h = C.c_void_p()
        rc = _lib().vdf_pipeline_create(self._ctx._h, max_batch_stacks, batch_bytes, _cropdetect_code(options.cropdetect),
                                        C.byref(self._h))
        if rc != 0:
            raise _ffi.VdfError(rc, "vdf_pipeline_create")

    def push(self, tag: int, frames: Sequence[np.ndarray]) -> None:
        """frames = the gray u8 frames of one video (any number; the first 16 are used, video_hash_builder.rs:164)"""
        fr = [np.ascontiguousarray(f, dtype=np.uint8) for f in list(frames)[:DCT_SIZE]]
        flags, w, h = 0, 0, 0
        if fr:
            h, w = fr[0].shape
            if any(f.shape != (h, w) for f in fr):  # are_all_frames_same_size, video_hash_builder.rs:169-186
                flags = _ffi.STACK_FLAG_MIXED_SIZES
        ptrs = (C.c_void_p * max(len(fr), 1))(*[f.ctypes.data for f in fr])
        rc = _lib().vdf_pipeline_push(self._h, int(tag), ptrs, len(fr), w, h, w, flags)
        if rc != 0:
            raise _ffi.VdfError(rc, "vdf_pipeline_push: " + _lib().vdf_pipeline_error(self._h).decode())

    def flush(self) -> None:
        rc = _lib().vdf_pipeline_flush(self._h)
        if rc != 0:
            raise _ffi.VdfError(rc, "vdf_pipeline_flush: " + _lib().vdf_pipeline_error(self._h).decode())

    def results(self, wait: bool = False, max_results: int = 4096) -> List[Tuple[int, Union[np.ndarray, Error], Tuple[int, int, int, int]]]:
        """-> [(tag, 16 hash words | the Error the reference would return, crop)]"""
        buf = (_Result * max_results)()
        n = C.c_uint32()
        rc = _lib().vdf_pipeline_poll(self._h, buf, max_results, C.byref(n), 1 if wait else 0)
        if rc != 0:
            raise _ffi.VdfError(rc, "vdf_pipeline_poll: " + _lib().vdf_pipeline_error(self._h).decode())
        out = []
        for r in buf[:n.value]:
            if r.status == _ffi.STACK_OK:
                val: Union[np.ndarray, Error] = np.array(r.hash[:], dtype=np.uint64)
            elif r.status == _ffi.STACK_VIDPROC:
                val = VidProc("frames not all same size")
            elif r.status == _ffi.STACK_NOT_ENOUGH_FRAMES:
                val = NotEnoughFrames()
            else:
                raise _ffi.VdfError(r.status, _lib().vdf_pipeline_error(self._h).decode())
            out.append((int(r.tag), val, tuple(r.crop[:])))
        return out

    def close(self) -> None:
        if self._h:
            _lib().vdf_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        try:
            if exc[0] is None:
                self.flush()
        finally:
            self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
