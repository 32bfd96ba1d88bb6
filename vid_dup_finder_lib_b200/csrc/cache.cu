// cache.cu -- reader / writer of the application's on-disk hash cache, host code only.
//
// The app keeps every VideoHash it ever computed in one file and loads it whole before a search
// (vid_dup_finder_app/src/video_hash_filesystem_cache/generic_filesystem_cache/base_fs_cache.rs:26,106-112,192-196):
//     HashMap<PathBuf, MtimeCacheEntry<Result<VideoHash, Error>>>        (processing_fs_cache.rs:23-27,
//                                                                         generic_cache_if.rs:22-23)
// serialised by bincode 2 through serde with `bincode::config::standard()` = little-endian, variable-length integers:
//     varint(u)        u < 251: one byte | 251 + u16 | 252 + u32 | 253 + u64           (every integer width, incl. usize)
//     map              varint(len), then (key, value) pairs in the map's iteration order (arbitrary for a HashMap)
//     PathBuf / String varint(byte length) + UTF-8 bytes                                (serde: Path -> str)
//     MtimeCacheEntry  cache_mtime, value                                               (struct = fields in order)
//     SystemTime       varint(secs_since_epoch: u64), varint(nanos_since_epoch: u32)    (serde's impl for SystemTime)
//     Result<T, E>     varint(0) + T  |  varint(1) + E                                  (serde: enum Result { Ok, Err })
//     VideoHash        16 x varint(usize word), src_path: string, varint(duration: u32) (video_hash.rs:26-32; arrays
//                                                                                        <= 32 are tuples: no length)
//     Error            varint(0) NotVideo | varint(1) + string VidProc | varint(2) NotEnoughFrames
//                                                                                       (video_hashing/mod.rs:17-28)
// bincode and serde are third-party crates that are not under /root/reference and no cache file is checked in, so this
// restates their published encodings: PARITY UNPINNED (tests hold the reader and writer against an independent Python
// encoder of the same rules and against each other).  Loading straight into struct-of-arrays is what lets a 10 M-hash
// corpus go from disk to vdf_search without ever existing as 10 M heap objects.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vdf_b200.h"

namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    uint64_t varint() {
        if (p >= end) return fail();
        const uint8_t b = *p++;
        if (b < 251) return b;
        const int n = b == 251 ? 2 : b == 252 ? 4 : b == 253 ? 8 : -1;
        if (n < 0 || end - p < n) return fail();  // 254 = u128: never written for these types
        uint64_t v = 0;
        for (int k = 0; k < n; ++k) v |= (uint64_t)p[k] << (8 * k);
        p += n;
        return v;
    }
    bool bytes(uint64_t n, const uint8_t** out) {
        if ((uint64_t)(end - p) < n) return fail() != 0;
        *out = p;
        p += n;
        return true;
    }
    uint64_t fail() {
        ok = false;
        p = end;
        return 0;
    }
};

struct Writer {
    std::vector<uint8_t> buf;
    void varint(uint64_t v) {
        if (v < 251) {
            buf.push_back((uint8_t)v);
            return;
        }
        const int n = v < (1ull << 16) ? 2 : v < (1ull << 32) ? 4 : 8;
        buf.push_back(n == 2 ? 251 : n == 4 ? 252 : 253);
        for (int k = 0; k < n; ++k) buf.push_back((uint8_t)(v >> (8 * k)));
    }
    void str(const char* s, uint64_t n) {
        varint(n);
        buf.insert(buf.end(), (const uint8_t*)s, (const uint8_t*)s + n);
    }
};

template <typename T>
T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc((v.size() ? v.size() : 1) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" {

void vdf_free_cache(vdf_cache* c) {
    if (!c) return;
    free(c->kind), free(c->hashes), free(c->durations), free(c->key_blob), free(c->key_off), free(c->src_blob);
    free(c->src_off), free(c->msg_blob), free(c->msg_off), free(c->mtime_secs), free(c->mtime_nanos);
    memset(c, 0, sizeof *c);
}

int vdf_cache_load(const char* file, vdf_cache* out) {
    if (!file || !out) return VDF_ERR_INVALID;
    memset(out, 0, sizeof *out);
    FILE* f = fopen(file, "rb");
    if (!f) return VDF_ERR_IO;
    std::vector<uint8_t> data;
    uint8_t chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) data.insert(data.end(), chunk, chunk + got);
    fclose(f);
    Reader r{data.data(), data.data() + data.size()};
    const uint64_t n = r.varint();
    if (!r.ok || n > data.size()) return VDF_ERR_FORMAT;  // every entry takes more than one byte
    std::vector<int32_t> kind(n);
    std::vector<uint64_t> hashes(n * 16, 0), key_off(n + 1, 0), src_off(n + 1, 0), msg_off(n + 1, 0), secs(n);
    std::vector<uint32_t> dur(n, 0), nanos(n);
    std::vector<char> keys, srcs, msgs;
    const uint8_t* s;
    for (uint64_t i = 0; i < n && r.ok; ++i) {
        uint64_t len = r.varint();  // key: PathBuf
        if (!r.bytes(len, &s)) break;
        keys.insert(keys.end(), s, s + len);
        secs[i] = r.varint();  // MtimeCacheEntry.cache_mtime
        const uint64_t ns = r.varint();
        if (ns >= 1000000000ull) r.fail();
        nanos[i] = (uint32_t)ns;
        const uint64_t variant = r.varint();  // MtimeCacheEntry.value: Result<VideoHash, Error>
        if (variant == 0) {
            kind[i] = VDF_CACHE_OK;
            for (int w = 0; w < 16; ++w) hashes[i * 16 + w] = r.varint();
            len = r.varint();
            if (!r.bytes(len, &s)) break;
            srcs.insert(srcs.end(), s, s + len);
            const uint64_t d = r.varint();
            if (d > 0xFFFFFFFFull) r.fail();
            dur[i] = (uint32_t)d;
        } else if (variant == 1) {
            const uint64_t e = r.varint();
            if (e == 0) kind[i] = VDF_CACHE_ERR_NOT_VIDEO;
            else if (e == 2) kind[i] = VDF_CACHE_ERR_NOT_ENOUGH_FRAMES;
            else if (e == 1) {
                kind[i] = VDF_CACHE_ERR_VIDPROC;
                len = r.varint();
                if (!r.bytes(len, &s)) break;
                msgs.insert(msgs.end(), s, s + len);
            } else r.fail();
        } else r.fail();
        key_off[i + 1] = keys.size(), src_off[i + 1] = srcs.size(), msg_off[i + 1] = msgs.size();
    }
    if (!r.ok || r.p != r.end) return VDF_ERR_FORMAT;  // truncated, malformed or trailing bytes
    out->n = n;
    out->kind = dup(kind), out->hashes = dup(hashes), out->durations = dup(dur);
    out->key_blob = dup(keys), out->key_off = dup(key_off), out->src_blob = dup(srcs), out->src_off = dup(src_off);
    out->msg_blob = dup(msgs), out->msg_off = dup(msg_off), out->mtime_secs = dup(secs), out->mtime_nanos = dup(nanos);
    if (!out->kind || !out->hashes || !out->durations || !out->key_blob || !out->key_off || !out->src_blob || !out->src_off ||
        !out->msg_blob || !out->msg_off || !out->mtime_secs || !out->mtime_nanos) {
        vdf_free_cache(out);
        return VDF_ERR_ALLOC;
    }
    return VDF_OK;
}

int vdf_cache_save(const char* file, const vdf_cache* c) {
    if (!file || !c) return VDF_ERR_INVALID;
    Writer w;
    w.varint(c->n);
    for (uint64_t i = 0; i < c->n; ++i) {
        w.str(c->key_blob + c->key_off[i], c->key_off[i + 1] - c->key_off[i]);
        w.varint(c->mtime_secs[i]);
        w.varint(c->mtime_nanos[i]);
        if (c->kind[i] == VDF_CACHE_OK) {
            w.varint(0);
            for (int k = 0; k < 16; ++k) w.varint(c->hashes[i * 16 + k]);
            w.str(c->src_blob + c->src_off[i], c->src_off[i + 1] - c->src_off[i]);
            w.varint(c->durations[i]);
        } else {
            w.varint(1);
            if (c->kind[i] == VDF_CACHE_ERR_NOT_VIDEO) w.varint(0);
            else if (c->kind[i] == VDF_CACHE_ERR_NOT_ENOUGH_FRAMES) w.varint(2);
            else if (c->kind[i] == VDF_CACHE_ERR_VIDPROC) {
                w.varint(1);
                w.str(c->msg_blob + c->msg_off[i], c->msg_off[i + 1] - c->msg_off[i]);
            } else return VDF_ERR_INVALID;
        }
    }
    // the reference writes a temporary file next to the cache and renames it over it (base_fs_cache.rs:84,157)
    const std::string tmp = std::string(file) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return VDF_ERR_IO;
    const bool wrote = fwrite(w.buf.data(), 1, w.buf.size(), f) == w.buf.size();
    if (fclose(f) != 0 || !wrote || rename(tmp.c_str(), file) != 0) {
        remove(tmp.c_str());
        return VDF_ERR_IO;
    }
    return VDF_OK;
}

}  // extern "C"
