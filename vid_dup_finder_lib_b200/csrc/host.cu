// host.cu -- the host half of `search` / `search_with_references` behind the C ABI (vdf_search, vdf_search_with_references,
// vdf_sort_order in include/vdf_b200.h): what video_dup_finder.rs:7-46 does around the comparison loops -- the stable
// (duration, src_path) sort of Search::sort (search_algorithm.rs:55-61, Rust `Path` ordering), the tolerance cast
// (search_algorithm.rs:82) and the mapping of sorted positions back to the caller's entries -- written natively so that a
// caller hands over its hashes in ITS order (struct-of-arrays, paths as one byte blob) and gets groups of ITS indices back.
// Multi-threaded (std::thread): the sort is the only O(n log n) host step on the path and at 1 M entries it costs more
// than the GPU's all-pairs comparison when done on one core.  No comparison work happens here: there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace {

using vdf::DevBuf;

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

unsigned n_threads(uint64_t n) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    hw = std::min(hw, 32u);
    const uint64_t by_work = n / 8192 + 1;  // do not spawn threads for small inputs
    return (unsigned)std::min<uint64_t>(hw, by_work);
}

template <typename F>
void parallel_for(uint64_t n, unsigned t, F&& body) {  // body(begin, end, thread)
    if (t <= 1 || n == 0) {
        body(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(t);
    for (unsigned k = 0; k < t; ++k) {
        const uint64_t b = n * k / t, e = n * (k + 1) / t;
        th.emplace_back([&body, b, e, k] { body(b, e, k); });
    }
    for (auto& x : th) x.join();
}

// ---- Rust `Path` ordering on Unix -------------------------------------------------------------------------------
// std::path::Path::cmp compares the component lists lexicographically; Component's derived order is
// RootDir < CurDir < ParentDir < Normal(bytes).  components(): a leading '/' is RootDir, empty pieces and interior "."
// are dropped, a leading "." of a relative path is CurDir, ".." is ParentDir.  Encoding every component as
// tag (1..4) + bytes + 0x00 makes plain byte order of the encodings equal that list order (no path byte is 0x00; a
// component that is a prefix of another ends in 0x00 < any byte; a list that is a prefix of another is shorter).
size_t encode_path(const char* p, size_t len, uint8_t* out) {  // out == nullptr: size only
    size_t w = 0;
    auto put = [&](uint8_t b) {
        if (out) out[w] = b;
        ++w;
    };
    size_t i = 0;
    const bool absolute = len > 0 && p[0] == '/';
    if (absolute) put(1), put(0);
    bool first_piece = true;
    while (i <= len) {
        size_t j = i;
        while (j < len && p[j] != '/') ++j;
        const size_t l = j - i;
        if (l == 0) {
            // empty piece (leading, doubled or trailing separator): dropped
        } else if (l == 1 && p[i] == '.') {
            if (first_piece && !absolute) put(2), put(0);
        } else if (l == 2 && p[i] == '.' && p[i + 1] == '.') {
            put(3), put(0);
        } else {
            put(4);
            for (size_t k = i; k < j; ++k) put((uint8_t)p[k]);
            put(0);
        }
        first_piece = false;
        i = j + 1;
    }
    return w;
}

struct SortKey {
    uint32_t dur;
    uint32_t idx;
    uint64_t prefix, prefix2;  // encoded bytes 0-7 and 8-15, big-endian, zero padded
};
constexpr uint64_t kPrefixBytes = 16;

struct KeyLess {
    const uint8_t* enc;
    const uint64_t* eoff;
    uint64_t skip;  // bytes shared by ALL encoded paths (their common directory prefix): never compared
    bool operator()(const SortKey& a, const SortKey& b) const {
        if (a.dur != b.dur) return a.dur < b.dur;
        if (a.prefix != b.prefix) return a.prefix < b.prefix;
        if (a.prefix2 != b.prefix2) return a.prefix2 < b.prefix2;
        const uint64_t la = eoff[a.idx + 1] - eoff[a.idx] - skip, lb = eoff[b.idx + 1] - eoff[b.idx] - skip;
        if (la > kPrefixBytes || lb > kPrefixBytes) {
            const uint64_t sa = la > kPrefixBytes ? la - kPrefixBytes : 0, sb = lb > kPrefixBytes ? lb - kPrefixBytes : 0;
            const int c = memcmp(enc + eoff[a.idx] + skip + (la > kPrefixBytes ? kPrefixBytes : la),
                                 enc + eoff[b.idx] + skip + (lb > kPrefixBytes ? kPrefixBytes : lb), std::min(sa, sb));
            if (c) return c < 0;
            if (sa != sb) return sa < sb;
        } else if (la != lb) {
            return la < lb;  // equal padded prefixes, different lengths: the shorter one is a prefix of the longer
        }
        return a.idx < b.idx;  // stable
    }
};

int sort_order_impl(const uint32_t* dur, const char* paths, const uint64_t* off, uint64_t n, std::vector<SortKey>& keys) {
    keys.resize(n);
    if (n == 0) return VDF_OK;
    const bool dbg = getenv("VDF_DEBUG_TIMING") != nullptr;
    double d0 = now_ms();
    const unsigned t = n_threads(n);
    std::vector<uint64_t> eoff(n + 1, 0);
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t i = b; i < e; ++i) eoff[i + 1] = encode_path(paths + off[i], off[i + 1] - off[i], nullptr);
    });
    for (uint64_t i = 0; i < n; ++i) eoff[i + 1] += eoff[i];
    std::vector<uint8_t> enc(eoff[n] + 8, 0);
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t i = b; i < e; ++i) encode_path(paths + off[i], off[i + 1] - off[i], enc.data() + eoff[i]);
    });
    // the bytes every path shares (a library usually lives under one directory) carry no order: the in-key prefixes start
    // after them, so they hold the bytes that actually tell paths apart
    std::vector<uint64_t> lcp(t, ~0ull);
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned th) {
        uint64_t m = eoff[1] - eoff[0];
        for (uint64_t i = b; i < e && m; ++i) {
            const uint64_t l = std::min(m, eoff[i + 1] - eoff[i]);
            uint64_t k = 0;
            while (k < l && enc[eoff[i] + k] == enc[k]) ++k;
            m = k;
        }
        lcp[th] = m;
    });
    uint64_t skip = eoff[1] - eoff[0];
    for (unsigned k = 0; k < t; ++k)
        if (lcp[k] != ~0ull) skip = std::min(skip, lcp[k]);
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t i = b; i < e; ++i) {
            const uint64_t l = eoff[i + 1] - eoff[i] - skip;
            const uint8_t* p = enc.data() + eoff[i] + skip;
            uint64_t pre = 0, pre2 = 0;
            for (uint64_t k = 0; k < 8; ++k) pre = (pre << 8) | (k < l ? p[k] : 0);
            for (uint64_t k = 8; k < 16; ++k) pre2 = (pre2 << 8) | (k < l ? p[k] : 0);
            keys[i] = SortKey{dur[i], (uint32_t)i, pre, pre2};
        }
    });
    KeyLess less{enc.data(), eoff.data(), skip};
    if (dbg) fprintf(stderr, "[vdf] sort: encode %.1f ms (%u threads)\n", now_ms() - d0, t), d0 = now_ms();
    if (t <= 1) {
        std::sort(keys.begin(), keys.end(), less);
        return VDF_OK;
    }
    // sorted runs in parallel, then pairwise merges (each round halves the number of runs)
    std::vector<uint64_t> cut(t + 1);
    for (unsigned k = 0; k <= t; ++k) cut[k] = n * k / t;
    parallel_for(t, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t k = b; k < e; ++k) std::sort(keys.begin() + cut[k], keys.begin() + cut[k + 1], less);
    });
    if (dbg) fprintf(stderr, "[vdf] sort: runs %.1f ms\n", now_ms() - d0), d0 = now_ms();
    // t sorted runs -> t disjoint key ranges (splitters from a regular sample of the runs), one thread per range: every
    // thread merges its t run segments on its own, so all merge passes run t-wide (pairwise merging of whole runs leaves the
    // last pass, over all n keys, to a single thread)
    std::vector<SortKey> sample;
    const uint64_t stride = std::max<uint64_t>(1, n / ((uint64_t)t * 64));
    for (unsigned r = 0; r < t; ++r)
        for (uint64_t i = cut[r] + stride / 2; i < cut[r + 1]; i += stride) sample.push_back(keys[i]);
    std::sort(sample.begin(), sample.end(), less);
    std::vector<uint64_t> pos((size_t)t * (t + 1));  // pos[r * (t + 1) + j]: first key of run r that belongs to range j or later
    for (unsigned r = 0; r < t; ++r) {
        pos[(size_t)r * (t + 1)] = cut[r];
        pos[(size_t)r * (t + 1) + t] = cut[r + 1];
    }
    parallel_for(t, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t r = b; r < e; ++r)
            for (unsigned j = 1; j < t; ++j) {
                const SortKey& sp = sample[std::min<size_t>(sample.size() - 1, sample.size() * j / t)];
                pos[r * (t + 1) + j] = (uint64_t)(std::lower_bound(keys.begin() + cut[r], keys.begin() + cut[r + 1], sp, less) - keys.begin());
            }
    });
    std::vector<uint64_t> out_off(t + 1, 0);
    for (unsigned j = 0; j < t; ++j) {
        uint64_t sz = 0;
        for (unsigned r = 0; r < t; ++r) sz += pos[(size_t)r * (t + 1) + j + 1] - pos[(size_t)r * (t + 1) + j];
        out_off[j + 1] = out_off[j] + sz;
    }
    std::vector<SortKey> tmp(n), tmp2(n);
    parallel_for(t, t, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t j = b; j < e; ++j) {
            // gather the t segments of range j back to back, then pairwise merges inside [out_off[j], out_off[j + 1])
            std::vector<uint64_t> seg(1, out_off[j]);
            uint64_t w = out_off[j];
            for (unsigned r = 0; r < t; ++r) {
                const uint64_t lo = pos[(size_t)r * (t + 1) + j], hi = pos[(size_t)r * (t + 1) + j + 1];
                std::copy(keys.begin() + lo, keys.begin() + hi, tmp.begin() + w);
                w += hi - lo;
                seg.push_back(w);
            }
            std::vector<SortKey>*src = &tmp, *dst = &tmp2;
            while (seg.size() > 2) {
                std::vector<uint64_t> next;
                for (size_t k = 0; k + 1 < seg.size(); k += 2) {
                    const uint64_t lo = seg[k], mid = seg[k + 1], hi = seg[std::min(k + 2, seg.size() - 1)];
                    std::merge(src->begin() + lo, src->begin() + mid, src->begin() + mid, src->begin() + hi, dst->begin() + lo, less);
                    next.push_back(lo);
                }
                next.push_back(seg.back());
                std::swap(src, dst);
                seg.swap(next);
            }
            if (src != &tmp) std::copy(src->begin() + out_off[j], src->begin() + out_off[j + 1], tmp.begin() + out_off[j]);
        }
    });
    keys.swap(tmp);
    if (dbg) fprintf(stderr, "[vdf] sort: merges %.1f ms\n", now_ms() - d0);
    return VDF_OK;
}

// (tolerance * 1000.0) as u32: truncating, saturating, NaN -> 0 (search_algorithm.rs:82)
uint32_t tolerance_to_int(double tolerance) {
    const double v = tolerance * 1000.0;
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

int enter(vdf_ctx* ctx) {
    if (!ctx) return VDF_ERR_INVALID;
    ctx->err.clear();
    if (cudaSetDevice(ctx->device) != cudaSuccess) {
        ctx->err = "cudaSetDevice failed";
        return VDF_ERR_CUDA;
    }
    return VDF_OK;
}

// hashes / durations in sorted order -> pinned staging -> HBM (async on the context's stream)
int stage_sorted(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* dur, const std::vector<SortKey>& keys, DevBuf& d_hash,
                 DevBuf& d_dur, vdf::PinnedBuf& pin_h, vdf::PinnedBuf& pin_d) {
    const uint64_t n = keys.size();
    VDF_ALLOC(ctx, pin_h.ensure(n * 128));
    VDF_ALLOC(ctx, pin_d.ensure(n * 4));
    VDF_ALLOC(ctx, d_hash.ensure(n * 128));
    VDF_ALLOC(ctx, d_dur.ensure(n * 4));
    uint64_t* ph = pin_h.as<uint64_t>();
    uint32_t* pd = pin_d.as<uint32_t>();
    parallel_for(n, n_threads(n), [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t k = b; k < e; ++k) {
            memcpy(ph + k * 16, hashes + (uint64_t)keys[k].idx * 16, 128);
            pd[k] = keys[k].dur;
        }
    });
    VDF_CUDA(ctx, cudaMemcpyAsync(d_hash.p, ph, n * 128, cudaMemcpyHostToDevice, ctx->stream));
    VDF_CUDA(ctx, cudaMemcpyAsync(d_dur.p, pd, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += n * 132;
    return VDF_OK;
}

template <typename F>
int with_growing_keys(vdf_ctx* ctx, F&& run, uint64_t* n_out) {
    uint64_t cap = ctx->initial_edges;
    for (int attempt = 0; attempt < 3; ++attempt) {
        VDF_ALLOC(ctx, ctx->keys_a.ensure(cap * 8));
        uint64_t cnt = 0;
        int rc = run(ctx->keys_a.as<uint64_t>(), cap, &cnt);
        if (rc == VDF_OK) {
            *n_out = cnt;
            return VDF_OK;
        }
        if (rc != VDF_ERR_EDGE_OVERFLOW) return rc;
        if (cnt > ctx->max_edges) {
            ctx->err = "edge buffer overflow: " + std::to_string(cnt) + " matches exceed max_edges " + std::to_string(ctx->max_edges);
            return VDF_ERR_EDGE_OVERFLOW;
        }
        cap = cnt + cnt / 16 + 1024;
    }
    return VDF_ERR_EDGE_OVERFLOW;
}

}  // namespace

extern "C" {

int vdf_sort_order(const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n, uint64_t* order_out) {
    if (n && (!durations || !path_off || !order_out || (!path_blob && path_off[n] != 0))) return VDF_ERR_INVALID;
    if (n >= 0xFFFFFFFFull) return VDF_ERR_INVALID;
    std::vector<SortKey> keys;
    int rc = sort_order_impl(durations, path_blob, path_off, n, keys);
    if (rc != VDF_OK) return rc;
    for (uint64_t k = 0; k < n; ++k) order_out[k] = keys[k].idx;
    return VDF_OK;
}

int vdf_stage_sorted(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob, const uint64_t* path_off,
                     uint64_t n, uint64_t* order_out, uint64_t* d_hash_dst, uint32_t* d_dur_dst, const uint64_t** d_hash_sorted,
                     const uint32_t** d_dur_sorted) {
    VDF_TRY(enter(ctx));
    if (!order_out || !d_hash_sorted || !d_dur_sorted || (n && (!hashes || !durations || !path_off)) || (!d_hash_dst != !d_dur_dst))
        return VDF_ERR_INVALID;
    if (n >= 0xFFFFFF00ull) {
        ctx->err = "n must be < 2^32";
        return VDF_ERR_INVALID;
    }
    double t0 = now_ms();
    std::vector<SortKey> keys;
    VDF_TRY(sort_order_impl(durations, path_blob, path_off, n, keys));
    double t1 = now_ms();
    if (n) VDF_TRY(stage_sorted(ctx, hashes, durations, keys, ctx->in_hash, ctx->in_dur, ctx->pin_a, ctx->pin_b));
    *d_hash_sorted = ctx->in_hash.as<uint64_t>();
    *d_dur_sorted = ctx->in_dur.as<uint32_t>();
    if (n && d_hash_dst) {  // the caller's buffers (e.g. tensors it will broadcast to the other ranks)
        VDF_CUDA(ctx, cudaMemcpyAsync(d_hash_dst, ctx->in_hash.p, n * 128, cudaMemcpyDeviceToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaMemcpyAsync(d_dur_dst, ctx->in_dur.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        *d_hash_sorted = d_hash_dst;
        *d_dur_sorted = d_dur_dst;
    }
    parallel_for(n, n_threads(n), [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t k = b; k < e; ++k) order_out[k] = keys[k].idx;
    });
    double t2 = now_ms();
    ctx->phase_ms[0] = t1 - t0, ctx->phase_ms[1] = t2 - t1, ctx->phase_ms[2] = 0, ctx->phase_ms[3] = 0;
    return VDF_OK;
}

int vdf_search(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob, const uint64_t* path_off,
               uint64_t n, double tolerance, vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out || (n && (!hashes || !durations || !path_off))) return VDF_ERR_INVALID;
    if (n >= 0xFFFFFF00ull) {
        ctx->err = "n must be < 2^32";
        return VDF_ERR_INVALID;
    }
    double t0 = now_ms();
    std::vector<SortKey> keys;
    VDF_TRY(sort_order_impl(durations, path_blob, path_off, n, keys));
    double t1 = now_ms();
    uint64_t ne = 0;
    if (n) {
        VDF_TRY(stage_sorted(ctx, hashes, durations, keys, ctx->in_hash, ctx->in_dur, ctx->pin_a, ctx->pin_b));
    }
    double t2 = now_ms();
    if (n) {
        const uint32_t tol_int = tolerance_to_int(tolerance);
        VDF_TRY(with_growing_keys(
            ctx,
            [&](uint64_t* k, uint64_t cap, uint64_t* cnt) {
                return vdf::search_self_device(ctx, ctx->in_hash.as<uint64_t>(), ctx->in_dur.as<uint32_t>(), n, tol_int, k, cap, cnt);
            },
            &ne));
    }
    VDF_TRY(vdf::group_device(ctx, n, ctx->keys_a.as<uint64_t>(), ne, out));
    double t3 = now_ms();
    // sorted positions -> the caller's indices (what `entries[i].value.src_path()` resolves to in the reference)
    const uint64_t total = out->n_groups ? out->group_ptr[out->n_groups] : 0;
    for (uint64_t k = 0; k < total; ++k) out->member_idx[k] = keys[out->member_idx[k]].idx;
    double t4 = now_ms();
    ctx->phase_ms[0] = t1 - t0, ctx->phase_ms[1] = t2 - t1, ctx->phase_ms[2] = t3 - t2, ctx->phase_ms[3] = t4 - t3;
    return VDF_OK;
}

int vdf_search_with_references(vdf_ctx* ctx, const uint64_t* ref_hashes, const uint32_t* ref_durations, uint64_t n_ref,
                               const uint64_t* cand_hashes, const uint32_t* cand_durations, const char* cand_path_blob,
                               const uint64_t* cand_path_off, uint64_t n_cand, double tolerance, vdf_csr* out) {
    VDF_TRY(enter(ctx));
    if (!out || (n_ref && (!ref_hashes || !ref_durations)) || (n_cand && (!cand_hashes || !cand_durations || !cand_path_off)))
        return VDF_ERR_INVALID;
    if (n_cand >= 0xFFFFFF00ull || n_ref >= 0xFFFFFF00ull) {
        ctx->err = "indices must fit in 32 bits";
        return VDF_ERR_INVALID;
    }
    out->n_rows = n_ref;
    out->row_ptr = (uint64_t*)calloc(n_ref + 1, 8);
    out->col_idx = nullptr;
    if (!out->row_ptr) return VDF_ERR_ALLOC;
    double t0 = now_ms();
    std::vector<SortKey> keys;
    VDF_TRY(sort_order_impl(cand_durations, cand_path_blob, cand_path_off, n_cand, keys));
    double t1 = now_ms();
    uint64_t ne = 0;
    double t2 = t1;
    if (n_cand && n_ref) {
        VDF_TRY(stage_sorted(ctx, cand_hashes, cand_durations, keys, ctx->in_hash, ctx->in_dur, ctx->pin_a, ctx->pin_b));
        VDF_ALLOC(ctx, ctx->in_hash2.ensure(n_ref * 128));
        VDF_ALLOC(ctx, ctx->in_dur2.ensure(n_ref * 4));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_hash2.p, ref_hashes, n_ref * 128, cudaMemcpyHostToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_dur2.p, ref_durations, n_ref * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += n_ref * 132;
        t2 = now_ms();
        const uint32_t tol_int = tolerance_to_int(tolerance);
        VDF_TRY(with_growing_keys(
            ctx,
            [&](uint64_t* k, uint64_t cap, uint64_t* cnt) {
                return vdf::search_refs_device(ctx, ctx->in_hash.as<uint64_t>(), ctx->in_dur.as<uint32_t>(), n_cand, 0,
                                               ctx->in_hash2.as<uint64_t>(), ctx->in_dur2.as<uint32_t>(), n_ref, tol_int, k, cap, cnt);
            },
            &ne));
    }
    out->col_idx = (uint64_t*)malloc((ne ? ne : 1) * 8);
    if (!out->col_idx) return VDF_ERR_ALLOC;
    double t3 = now_ms();
    if (ne) {
        VDF_CUDA(ctx, cudaMemcpyAsync(out->col_idx, ctx->keys_a.p, ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2h += ne * 8;
        t3 = now_ms();
        // keys are sorted by (ref, sorted candidate position): count per ref, map positions to the caller's indices
        for (uint64_t k = 0; k < ne; ++k) {
            out->row_ptr[(out->col_idx[k] >> 32) + 1]++;
            out->col_idx[k] = keys[out->col_idx[k] & 0xFFFFFFFFull].idx;
        }
        for (uint64_t r = 0; r < n_ref; ++r) out->row_ptr[r + 1] += out->row_ptr[r];
    }
    double t4 = now_ms();
    ctx->phase_ms[0] = t1 - t0, ctx->phase_ms[1] = t2 - t1, ctx->phase_ms[2] = t3 - t2, ctx->phase_ms[3] = t4 - t3;
    return VDF_OK;
}

int vdf_ctx_last_phases(const vdf_ctx* ctx, double* ms4) {
    if (!ctx || !ms4) return VDF_ERR_INVALID;
    for (int k = 0; k < 4; ++k) ms4[k] = ctx->phase_ms[k];
    return VDF_OK;
}

}  // extern "C"
