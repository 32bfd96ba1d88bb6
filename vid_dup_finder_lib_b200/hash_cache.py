"""The application's on-disk hash cache: read / write the file the reference app keeps its
VideoHashes in, straight into / out of the struct-of-arrays `HashTable` the GPU search consumes.

Format (vid_dup_finder_app/src/video_hash_filesystem_cache/generic_filesystem_cache/base_fs_cache.rs:26,106-112,192-196):
bincode 2 `standard()` of HashMap<PathBuf, MtimeCacheEntry<Result<VideoHash, Error>>>; the decoding itself is native
(vdf_cache_load / vdf_cache_save, csrc/cache.cu, which states the encoding rule by rule).  Next to the cache lives
`<stem>.metadata.txt` (video_hash_filesystem_cache.rs:104, cache_metadata.rs:45-91): one CSV line
`operating_system,decode_backend,crop,skip_forward_amount,cache_version` written with `{:?},{:?},{:?},{},{}`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _ffi
from .definitions import Cropdetect
from .video_hash import HashTable
from .video_hash_builder import Error, NotEnoughFrames, NotVideo, VidProc

CACHE_VERSION = 1  # cache_metadata.rs:62


class _Cache(C.Structure):
    _fields_ = [("n", C.c_uint64), ("kind", C.POINTER(C.c_int32)), ("hashes", C.POINTER(C.c_uint64)),
                ("durations", C.POINTER(C.c_uint32)), ("key_blob", C.POINTER(C.c_char)), ("key_off", C.POINTER(C.c_uint64)),
                ("src_blob", C.POINTER(C.c_char)), ("src_off", C.POINTER(C.c_uint64)), ("msg_blob", C.POINTER(C.c_char)),
                ("msg_off", C.POINTER(C.c_uint64)), ("mtime_secs", C.POINTER(C.c_uint64)),
                ("mtime_nanos", C.POINTER(C.c_uint32))]


def _lib():
    L = _ffi.lib()
    L.vdf_cache_load.argtypes = [C.c_char_p, C.POINTER(_Cache)]
    L.vdf_cache_save.argtypes = [C.c_char_p, C.POINTER(_Cache)]
    L.vdf_free_cache.argtypes = [C.POINTER(_Cache)]
    L.vdf_free_cache.restype = None
    return L


def _arr(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)


def _split(blob: bytes, off: np.ndarray) -> List[str]:
    o = off.tolist()
    return [os.fsdecode(blob[a:b]) for a, b in zip(o[:-1], o[1:])]


@dataclass
class HashCache:
    """Every entry of a cache file, in file order.  kind: 0 = Ok(VideoHash), 1/2/3 = Err(NotVideo / VidProc / NotEnoughFrames)
    (include/vdf_b200.h VDF_CACHE_*)."""

    kind: np.ndarray         # int32 [n]
    hashes: np.ndarray       # uint64 [n,16]  (zero rows for error entries)
    durations: np.ndarray    # uint32 [n]
    keys: List[str]          # the map keys (file paths)
    src_paths: List[str]     # VideoHash.src_path ("" for error entries)
    messages: List[str]      # Error::VidProc's String ("" elsewhere)
    mtime_secs: np.ndarray   # uint64 [n]
    mtime_nanos: np.ndarray  # uint32 [n]

    def __len__(self):
        return len(self.kind)

    def table(self) -> HashTable:
        """the Ok entries as the search table (what `all_cached_items` + filter(Ok) feeds `search`, app_fns.rs:428-482)"""
        ok = np.nonzero(self.kind == 0)[0]
        return HashTable(self.hashes[ok], self.durations[ok], [self.src_paths[i] for i in ok.tolist()])

    def errors(self) -> Dict[str, Error]:
        ctor = {1: lambda m: NotVideo(), 2: lambda m: VidProc(m), 3: lambda m: NotEnoughFrames()}
        return {self.keys[i]: ctor[int(self.kind[i])](self.messages[i]) for i in np.nonzero(self.kind != 0)[0].tolist()}

    @staticmethod
    def from_table(table: HashTable, mtime_secs=0, mtime_nanos=0) -> "HashCache":
        n = len(table)
        return HashCache(np.zeros(n, np.int32), table.hashes.copy(), table.durations.copy(), list(table.paths), list(table.paths),
                         [""] * n, np.full(n, mtime_secs, np.uint64), np.full(n, mtime_nanos, np.uint32))


def load_hash_cache(path) -> HashCache:
    c = _Cache()
    rc = _lib().vdf_cache_load(os.fsencode(path), C.byref(c))
    if rc != 0:
        raise _ffi.VdfError(rc, f"cannot read hash cache {path!r}: " + {-6: "I/O error", -7: "not a bincode-2 hash cache"}.get(rc, ""))
    try:
        n = int(c.n)
        ko, so, mo = _arr(c.key_off, n + 1, np.uint64), _arr(c.src_off, n + 1, np.uint64), _arr(c.msg_off, n + 1, np.uint64)
        return HashCache(_arr(c.kind, n, np.int32), _arr(c.hashes, n * 16, np.uint64).reshape(n, 16), _arr(c.durations, n, np.uint32),
                         _split(C.string_at(c.key_blob, int(ko[-1])), ko), _split(C.string_at(c.src_blob, int(so[-1])), so),
                         _split(C.string_at(c.msg_blob, int(mo[-1])), mo), _arr(c.mtime_secs, n, np.uint64),
                         _arr(c.mtime_nanos, n, np.uint32))
    finally:
        _lib().vdf_free_cache(C.byref(c))


def save_hash_cache(path, cache: HashCache) -> None:
    n = len(cache)

    def blob(strs):
        enc = [os.fsencode(s) for s in strs]
        off = np.zeros(n + 1, np.uint64)
        if n:
            np.cumsum(np.fromiter(map(len, enc), np.uint64, n), out=off[1:])
        return C.create_string_buffer(b"".join(enc) + b"\0"), off

    kb, ko = blob(cache.keys)
    sb, so = blob(cache.src_paths)
    mb, mo = blob(cache.messages)
    kind = np.ascontiguousarray(cache.kind, np.int32)
    hashes = np.ascontiguousarray(cache.hashes, np.uint64).reshape(-1)
    dur = np.ascontiguousarray(cache.durations, np.uint32)
    secs = np.ascontiguousarray(cache.mtime_secs, np.uint64)
    nanos = np.ascontiguousarray(cache.mtime_nanos, np.uint32)

    def p(a, t):
        return a.ctypes.data_as(C.POINTER(t))

    c = _Cache(n, p(kind, C.c_int32), p(hashes, C.c_uint64), p(dur, C.c_uint32), C.cast(kb, C.POINTER(C.c_char)), p(ko, C.c_uint64),
               C.cast(sb, C.POINTER(C.c_char)), p(so, C.c_uint64), C.cast(mb, C.POINTER(C.c_char)), p(mo, C.c_uint64),
               p(secs, C.c_uint64), p(nanos, C.c_uint32))
    rc = _lib().vdf_cache_save(os.fsencode(path), C.byref(c))
    if rc != 0:
        raise _ffi.VdfError(rc, f"cannot write hash cache {path!r}")


# ---- <stem>.metadata.txt (cache_metadata.rs) ---------------------------------------------------------------------
@dataclass(frozen=True)
class CacheMetadata:
    operating_system: str = "Unix"            # cache_metadata.rs:6-10
    decode_backend: str = "FfmpegBackend"     # :25-29 (GstreamerBackend with the gstreamer_backend feature)
    crop: Cropdetect = Cropdetect.LETTERBOX
    skip_forward_amount: float = 15.0
    cache_version: int = CACHE_VERSION

    def to_disk_fmt(self) -> str:  # cache_metadata.rs:71-80 : "{:?},{:?},{:?},{},{}"
        f = self.skip_forward_amount
        amount = str(int(f)) if float(f).is_integer() and abs(f) < 1e16 else repr(float(f))  # Rust `{}` prints 15.0 as "15"
        return f"{self.operating_system},{self.decode_backend},{self.crop.name.capitalize()},{amount},{self.cache_version}"

    @staticmethod
    def try_parse(val: str) -> "CacheMetadata":  # cache_metadata.rs:82-122 (trim + lowercase on the enums)
        parts = val.split(",")
        if len(parts) != 5:
            raise ValueError(f"Could not parse cache metadata. Got {val}")
        osys = {"windows": "Windows", "unix": "Unix"}.get(parts[0].strip().lower())
        backend = {"ffmpegbackend": "FfmpegBackend", "gstreamerbackend": "GstreamerBackend"}.get(parts[1].strip().lower())
        crop = {c.name.lower(): c for c in Cropdetect}.get(parts[2].strip().lower())
        if osys is None or backend is None or crop is None:
            raise ValueError(f"Could not parse cache metadata. Got {val}")
        return CacheMetadata(osys, backend, crop, float(parts[3]), int(parts[4]))

    def validate(self, exp_crop: Cropdetect, exp_skip_forward_amount: float, exp_backend: str = "FfmpegBackend") -> Optional[str]:
        """cache_metadata.rs:124-162: the first mismatch, or None"""
        exp = CacheMetadata("Unix", exp_backend, exp_crop, exp_skip_forward_amount)
        for name in ("operating_system", "decode_backend", "crop", "skip_forward_amount", "cache_version"):
            if getattr(self, name) != getattr(exp, name):
                return f"{name} mismatch: Act: {getattr(self, name)!r}, Exp: {getattr(exp, name)!r}"
        return None


def metadata_path(cache_path) -> str:  # video_hash_filesystem_cache.rs:104
    d, base = os.path.split(os.fspath(cache_path))
    return os.path.join(d, os.path.splitext(base)[0] + ".metadata.txt")
