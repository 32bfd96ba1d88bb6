"""The app's hash-cache file: vdf_cache_load / vdf_cache_save (csrc/cache.cu) against an
independent statement of bincode 2 `standard()` + serde's encodings written here in Python (parity UNPINNED: no Rust
toolchain, no cache file in the reference tree; see the header of csrc/cache.cu for the rule-by-rule citation)."""
import os
import struct

import numpy as np
import pytest

import vid_dup_finder_lib_b200 as vdf
from vid_dup_finder_lib_b200 import _ffi, hash_cache as hc
from vid_dup_finder_lib_b200.definitions import Cropdetect


# ---- bincode 2, config::standard(): little-endian, varint -------------------------------------------------------
def varint(u: int) -> bytes:
    if u < 251:
        return bytes([u])
    if u < 1 << 16:
        return b"\xfb" + struct.pack("<H", u)
    if u < 1 << 32:
        return b"\xfc" + struct.pack("<I", u)
    return b"\xfd" + struct.pack("<Q", u)


def string(s: str) -> bytes:
    b = s.encode()
    return varint(len(b)) + b


def video_hash(words, src_path, duration) -> bytes:  # video_hash.rs:26-32: [usize; 16], PathBuf, u32
    return b"".join(varint(int(w)) for w in words) + string(src_path) + varint(duration)


def entry(key, secs, nanos, value) -> bytes:  # (PathBuf, MtimeCacheEntry { cache_mtime: SystemTime, value: Result<..> })
    out = string(key) + varint(secs) + varint(nanos)
    if value[0] == "ok":
        return out + varint(0) + video_hash(*value[1:])
    err = {"NotVideo": varint(0), "VidProc": varint(1) + string(value[2] if len(value) > 2 else ""), "NotEnoughFrames": varint(2)}
    return out + varint(1) + err[value[1]]


def cache_file(entries) -> bytes:  # HashMap: varint(len) + pairs
    return varint(len(entries)) + b"".join(entry(*e) for e in entries)


def _sample_entries():
    rng = np.random.default_rng(5)
    w1 = rng.integers(0, 2**63, 16, dtype=np.uint64) * 2 + 1  # 9-byte varints
    w2 = np.array([0, 1, 250, 251, 252, 65535, 65536, 2**32 - 1, 2**32, 2**64 - 1, 7, 8, 9, 10, 11, 12], dtype=np.uint64)  # every width
    return [
        ("/videos/cat.1.mp4", 1_700_000_000, 123_456_789, ("ok", w1, "/videos/cat.1.mp4", 29)),
        ("/videos/été 2019/dog 3.webm", 1_600_000_000, 0, ("ok", w2, "/videos/été 2019/dog 3.webm", 70000)),
        ("/videos/readme.txt", 5, 999_999_999, ("err", "NotVideo")),
        ("/videos/broken.mkv", 2**33, 1, ("err", "VidProc", "ffmpeg exited with ☠ status 1")),
        ("/videos/short.mp4", 250, 251, ("err", "NotEnoughFrames")),
        ("x" * 300, 0, 0, ("ok", np.zeros(16, np.uint64), "x" * 300, 0)),  # a 2-byte length varint
    ]


def test_varint_rule():
    assert varint(250) == b"\xfa" and varint(251) == b"\xfb\xfb\x00" and varint(65536) == b"\xfc\x00\x00\x01\x00"
    assert varint(2**32) == b"\xfd" + struct.pack("<Q", 2**32)


def test_load_reads_every_entry_kind(tmp_path):
    ents = _sample_entries()
    f = tmp_path / "vdf_cache.bin"
    f.write_bytes(cache_file(ents))
    c = hc.load_hash_cache(f)
    assert len(c) == len(ents) and c.kind.tolist() == [0, 0, 1, 2, 3, 0]
    assert c.keys == [e[0] for e in ents]
    assert c.mtime_secs.tolist() == [e[1] for e in ents] and c.mtime_nanos.tolist() == [e[2] for e in ents]
    for i in (0, 1, 5):
        assert np.array_equal(c.hashes[i], ents[i][3][1]) and c.src_paths[i] == ents[i][3][2] and c.durations[i] == ents[i][3][3]
    assert c.messages[3] == "ffmpeg exited with ☠ status 1"
    t = c.table()  # the Ok entries, ready for the GPU search
    assert isinstance(t, vdf.HashTable) and len(t) == 3 and t.paths[1].endswith("dog 3.webm") and t.durations.tolist() == [29, 70000, 0]
    errs = c.errors()
    assert set(errs) == {"/videos/readme.txt", "/videos/broken.mkv", "/videos/short.mp4"}
    assert str(errs["/videos/broken.mkv"]) == "Video processing error: ffmpeg exited with ☠ status 1"
    assert str(errs["/videos/readme.txt"]) == "File is not a video" and str(errs["/videos/short.mp4"]) == "Could not extract enough frames"


def test_save_writes_the_same_bytes_and_round_trips(tmp_path):
    ents = _sample_entries()
    want = cache_file(ents)
    f, g = tmp_path / "in.bin", tmp_path / "out.bin"
    f.write_bytes(want)
    hc.save_hash_cache(g, hc.load_hash_cache(f))
    assert g.read_bytes() == want and not os.path.exists(str(g) + ".tmp")
    # a table of generated hashes: save, load, identical
    rng = np.random.default_rng(6)
    n = 5000
    tb = vdf.HashTable(rng.integers(0, 2**64, (n, 16), dtype=np.uint64), rng.integers(0, 20000, n).astype(np.uint32),
                       ["/lib/%02d/v%06d.mkv" % (i % 37, i) for i in range(n)])
    hc.save_hash_cache(g, hc.HashCache.from_table(tb, 1_700_000_000, 5))
    back = hc.load_hash_cache(g).table()
    assert np.array_equal(back.hashes, tb.hashes) and np.array_equal(back.durations, tb.durations) and back.paths == tb.paths
    assert g.read_bytes() == cache_file([(p, 1_700_000_000, 5, ("ok", tb.hashes[i], p, int(tb.durations[i]))) for i, p in enumerate(tb.paths)])


def test_empty_truncated_and_garbage_files(tmp_path):
    f = tmp_path / "c.bin"
    f.write_bytes(cache_file([]))
    assert len(hc.load_hash_cache(f)) == 0 and len(hc.load_hash_cache(f).table()) == 0
    good = cache_file(_sample_entries())
    for bad in (good[:-1], good[:40], good + b"\0", b"", b"\xfe" + good[1:], varint(3) + good[1:]):
        f.write_bytes(bad)
        with pytest.raises(_ffi.VdfError):
            hc.load_hash_cache(f)
    with pytest.raises(_ffi.VdfError):
        hc.load_hash_cache(tmp_path / "missing.bin")


def test_cache_metadata_file():  # cache_metadata.rs:45-162, video_hash_filesystem_cache.rs:104
    m = hc.CacheMetadata(crop=Cropdetect.LETTERBOX, skip_forward_amount=15.0)
    assert m.to_disk_fmt() == "Unix,FfmpegBackend,Letterbox,15,1"
    assert hc.CacheMetadata(crop=Cropdetect.NONE, skip_forward_amount=7.5).to_disk_fmt() == "Unix,FfmpegBackend,None,7.5,1"
    assert hc.CacheMetadata.try_parse(" unix ,FFMPEGBACKEND,letterbox,15,1") == m
    assert m.validate(Cropdetect.LETTERBOX, 15.0) is None
    assert m.validate(Cropdetect.NONE, 15.0).startswith("crop mismatch") and m.validate(Cropdetect.LETTERBOX, 0.0).startswith("skip_forward_amount")
    for bad in ("unix,ffmpegbackend,letterbox,15", "beos,ffmpegbackend,letterbox,15,1", "unix,ffmpegbackend,letterbox,x,1"):
        with pytest.raises(ValueError):
            hc.CacheMetadata.try_parse(bad)
    assert hc.metadata_path("/a/b/vdf_cache.bin") == "/a/b/vdf_cache.metadata.txt"
