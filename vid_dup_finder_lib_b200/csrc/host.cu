// host.cu -- the host half of `search` / `search_with_references` behind the C ABI (vdf_search, vdf_search_with_references,
// vdf_sort_order in include/vdf_b200.h): what video_dup_finder.rs:7-46 does around the comparison loops -- the stable
// (duration, src_path) sort of Search::sort (search_algorithm.rs:55-61, Rust `Path` ordering), the tolerance cast
// (search_algorithm.rs:82) and the mapping of sorted positions back to the caller's entries -- written natively so that a
// caller hands over its hashes in ITS order (struct-of-arrays, paths as one byte blob) and gets groups of ITS indices back.
// The (duration, Path) sort is the only O(n log n) step around the comparison and at 1 M entries a host sort costs a third of
// the GPU's all-pairs time, so vdf_search / vdf_search_with_references / vdf_stage_sorted sort ON THE GPU: host threads only
// re-encode each path so that byte order equals Rust's component order and cut a 16-byte key after the bytes all paths share
// (two streaming passes, no per-path allocation); the GPU radix-sorts (duration, key, index) while the host is still copying
// the hashes into pinned memory; entries whose 20-byte keys tie (rare: paths that agree beyond their first 16 distinguishing
// bytes) are put in order on the host inside their runs only.  The table is uploaded in the CALLER's order and gathered by
// the sorted permutation when it is packed for the pair kernel, and group members are mapped back to the caller's indices
// on the device.  vdf_sort_order (no context, no GPU) keeps the multi-threaded host merge sort.
// No comparison work happens here: there is no CPU fallback.
#include <cub/device/device_radix_sort.cuh>

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

// encode_path, stopping once `cap` bytes are written: -> bytes written (<= cap)
size_t encode_path_capped(const char* p, size_t len, uint8_t* out, size_t cap) {
    size_t w = 0;
    auto put = [&](uint8_t b) {
        if (w < cap) out[w] = b;
        ++w;
    };
    size_t i = 0;
    const bool absolute = len > 0 && p[0] == '/';
    if (absolute) put(1), put(0);
    bool first_piece = true;
    while (i <= len && w < cap) {
        size_t j = i;
        while (j < len && p[j] != '/') ++j;
        const size_t l = j - i;
        if (l == 0) {
        } else if (l == 1 && p[i] == '.') {
            if (first_piece && !absolute) put(2), put(0);
        } else if (l == 2 && p[i] == '.' && p[i + 1] == '.') {
            put(3), put(0);
        } else {
            put(4);
            for (size_t k = i; k < j && w < cap; ++k) put((uint8_t)p[k]);
            put(0);
        }
        first_piece = false;
        i = j + 1;
    }
    return std::min(w, cap);
}

// (duration, 16 encoded bytes after the common prefix) per entry, struct-of-arrays in pinned memory:
// pd[n] u32, p1[n] u64, p2[n] u64 (big-endian, zero padded).  Two streaming passes over the paths.
void build_prefix_keys(const uint32_t* dur, const char* paths, const uint64_t* off, uint64_t n, uint32_t* pd, uint64_t* p1, uint64_t* p2,
                       uint64_t* skip_out) {
    const unsigned t = n_threads(n);
    std::vector<uint8_t> e0(encode_path(paths + off[0], off[1] - off[0], nullptr));
    encode_path(paths + off[0], off[1] - off[0], e0.data());
    std::vector<uint64_t> lcp(t, e0.size());
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned th) {
        uint64_t m = e0.size();
        std::vector<uint8_t> buf(e0.size() + 1);
        for (uint64_t i = b; i < e && m; ++i) {
            const size_t w = encode_path_capped(paths + off[i], off[i + 1] - off[i], buf.data(), m);
            uint64_t k = 0;
            while (k < w && buf[k] == e0[k]) ++k;
            m = k;
        }
        lcp[th] = m;
    });
    uint64_t skip = e0.size();
    for (unsigned k = 0; k < t; ++k) skip = std::min(skip, lcp[k]);
    if (n == 1) skip = 0;
    parallel_for(n, t, [&](uint64_t b, uint64_t e, unsigned) {
        std::vector<uint8_t> buf(skip + kPrefixBytes);
        for (uint64_t i = b; i < e; ++i) {
            const size_t w = encode_path_capped(paths + off[i], off[i + 1] - off[i], buf.data(), skip + kPrefixBytes);
            const uint64_t l = w > skip ? w - skip : 0;
            const uint8_t* p = buf.data() + skip;
            uint64_t a = 0, c = 0;
            for (uint64_t k = 0; k < 8; ++k) a = (a << 8) | (k < l ? p[k] : 0);
            for (uint64_t k = 8; k < 16; ++k) c = (c << 8) | (k < l ? p[k] : 0);
            pd[i] = dur[i], p1[i] = a, p2[i] = c;
        }
    });
    *skip_out = skip;
}

__global__ void iota_kernel(uint32_t* __restrict__ a, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}
template <typename T>
__global__ void gather_kernel(const T* __restrict__ in, const uint32_t* __restrict__ idx, uint64_t n, T* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}
// adjacent entries of the sorted order whose whole 20-byte keys are equal: their order is not decided yet
__global__ void tie_count_kernel(const uint32_t* __restrict__ ds, const uint64_t* __restrict__ p1, const uint64_t* __restrict__ p2,
                                 const uint32_t* __restrict__ order, uint64_t n, unsigned long long* __restrict__ ties) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k == 0 || k >= n) return;
    const uint32_t a = order[k - 1], b = order[k];
    if (ds[k - 1] == ds[k] && p1[a] == p1[b] && p2[a] == p2[b]) atomicAdd(ties, 1ull);
}
__global__ void widen_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// Search::sort (search_algorithm.rs:55-61) on the GPU.  In: the pinned key arrays of build_prefix_keys.  Out, in HBM:
// ctx->sk_order[k] = the caller's index of the k-th sorted entry, ctx->in_dur = the durations in sorted order.
// Three stable LSD passes (key bytes 8-15, key bytes 0-7, duration) carry the permutation; equal keys keep index order.
int gpu_sort(vdf_ctx* ctx, const uint32_t* dur, const char* paths, const uint64_t* off, uint64_t n, const uint32_t* pd, const uint64_t* p1,
             const uint64_t* p2, uint64_t skip) {
    cudaStream_t st = ctx->stream;
    const unsigned B = 256, G = (unsigned)((n + B - 1) / B);
    VDF_ALLOC(ctx, ctx->sk_a.ensure(n * 4));
    VDF_ALLOC(ctx, ctx->sk_b.ensure(n * 8));
    VDF_ALLOC(ctx, ctx->sk_c.ensure(n * 8));
    VDF_ALLOC(ctx, ctx->sk_d.ensure(n * (8 + 8 + 4 + 4 + 4) + 64));
    VDF_ALLOC(ctx, ctx->sk_order.ensure(n * 4));
    VDF_ALLOC(ctx, ctx->in_dur.ensure(n * 4));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->sk_a.p, pd, n * 4, cudaMemcpyHostToDevice, st));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->sk_b.p, p1, n * 8, cudaMemcpyHostToDevice, st));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->sk_c.p, p2, n * 8, cudaMemcpyHostToDevice, st));
    ctx->h2d += n * 20;
    uint64_t* k0 = ctx->sk_d.as<uint64_t>();
    uint64_t* k1 = k0 + n;
    uint32_t* ia = reinterpret_cast<uint32_t*>(k1 + n);
    uint32_t* ib = ia + n;
    uint32_t* d0 = ib + n;
    unsigned long long* ties = reinterpret_cast<unsigned long long*>(ctx->sk_d.as<uint8_t>() + ((n * 28 + 7) / 8) * 8);
    const uint32_t* d_dur = ctx->sk_a.as<uint32_t>();
    const uint64_t* d_p1 = ctx->sk_b.as<uint64_t>();
    const uint64_t* d_p2 = ctx->sk_c.as<uint64_t>();
    uint32_t* order = ctx->sk_order.as<uint32_t>();
    uint32_t* ds = ctx->in_dur.as<uint32_t>();
    size_t tmp = 0, tmp32 = 0;
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_p2, k1, ia, ib, (size_t)n, 0, 64, st));
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp32, d0, ds, ia, order, (size_t)n, 0, 32, st));
    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(std::max(tmp, tmp32)));
    VDF_CUDA(ctx, cudaMemsetAsync(ties, 0, 8, st));
    iota_kernel<<<G, B, 0, st>>>(ia, n);
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp, d_p2, k1, ia, ib, (size_t)n, 0, 64, st));
    gather_kernel<uint64_t><<<G, B, 0, st>>>(d_p1, ib, n, k0);
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp, k0, k1, ib, ia, (size_t)n, 0, 64, st));
    gather_kernel<uint32_t><<<G, B, 0, st>>>(d_dur, ia, n, d0);
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp32, d0, ds, ia, order, (size_t)n, 0, 32, st));
    tie_count_kernel<<<G, B, 0, st>>>(ds, d_p1, d_p2, order, n, ties);
    ctx->launches += 4 + 3 * 4;
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        ctx->err = std::string("gpu_sort: ") + cudaGetErrorString(le);
        return VDF_ERR_CUDA;
    }
    VDF_ALLOC(ctx, ctx->h_misc.ensure(256));
    unsigned long long* h = ctx->h_misc.as<unsigned long long>();
    VDF_CUDA(ctx, cudaMemcpyAsync(h + 16, ties, 8, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    if (h[16] == 0) return VDF_OK;
    // Ties: bring the order back and finish the comparison inside every run of equal keys with the full encodings.  The GPU
    // sort is stable, so a run lists its entries by ascending index, and a stable sort by the encoded bytes completes it.
    std::vector<uint32_t> ord(n);
    VDF_CUDA(ctx, cudaMemcpyAsync(ord.data(), order, n * 4, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->d2h += n * 4;
    auto same = [&](uint32_t a, uint32_t b) { return pd[a] == pd[b] && p1[a] == p1[b] && p2[a] == p2[b]; };
    parallel_for(n, n_threads(n), [&](uint64_t b, uint64_t e, unsigned) {
        std::vector<std::pair<std::string, uint32_t>> run;
        uint64_t k = b;
        while (k < e) {
            if (k > 0 && same(ord[k - 1], ord[k])) {  // inside a run that an earlier range owns
                ++k;
                continue;
            }
            uint64_t r = k + 1;
            while (r < n && same(ord[k], ord[r])) ++r;
            if (r - k > 1) {
                run.clear();
                for (uint64_t q = k; q < r; ++q) {
                    const uint32_t i = ord[q];
                    std::string enc(encode_path(paths + off[i], off[i + 1] - off[i], nullptr), '\0');
                    encode_path(paths + off[i], off[i + 1] - off[i], reinterpret_cast<uint8_t*>(&enc[0]));
                    run.emplace_back(std::move(enc), i);
                }
                std::stable_sort(run.begin(), run.end(), [](const std::pair<std::string, uint32_t>& x, const std::pair<std::string, uint32_t>& y) {
                    const size_t m = std::min(x.first.size(), y.first.size());
                    const int c = memcmp(x.first.data(), y.first.data(), m);
                    return c ? c < 0 : x.first.size() < y.first.size();
                });
                for (uint64_t q = k; q < r; ++q) ord[q] = run[q - k].second;
            }
            k = r;
        }
    });
    (void)dur, (void)skip;
    VDF_CUDA(ctx, cudaMemcpyAsync(order, ord.data(), n * 4, cudaMemcpyHostToDevice, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));  // `ord` goes out of scope
    ctx->h2d += n * 4;
    return VDF_OK;
}

int enter(vdf_ctx* ctx) {
    if (!ctx) return VDF_ERR_INVALID;
    ctx->err.clear();
    if (cudaSetDevice(ctx->device) != cudaSuccess) {
        ctx->err = "cudaSetDevice failed";
        return VDF_ERR_CUDA;
    }
    cudaGetLastError();  // an error a previous call already reported must not be picked up by this one's launch checks
    return VDF_OK;
}

}  // namespace

// The caller's table -> HBM, sorted on the way: keys are cut on the host and sorted on the GPU (gpu_sort) while host threads
// copy the hashes, in the CALLER's order, through pinned memory (each thread enqueues the upload of its own slice, so the
// DMA overlaps the copying).  Out: d_hash = the unsorted hashes, ctx->sk_order = the sorted permutation, ctx->in_dur = the
// sorted durations.  t_keys / t_stage: host milliseconds of the two phases.
int vdf::stage_and_sort(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* dur, const char* paths, const uint64_t* off, uint64_t n,
                        DevBuf& d_hash, vdf::PinnedBuf& pin_h, double* t_keys, double* t_stage) {
    const double t0 = now_ms();
    VDF_ALLOC(ctx, ctx->pin_c.ensure(n * 20));
    VDF_ALLOC(ctx, pin_h.ensure(n * 128));
    VDF_ALLOC(ctx, d_hash.ensure(n * 128));
    uint64_t* p1 = ctx->pin_c.as<uint64_t>();
    uint64_t* p2 = p1 + n;
    uint32_t* pd = reinterpret_cast<uint32_t*>(p2 + n);
    // The hash upload starts first and runs BESIDE the key cutting: copier threads move the caller's (pageable) table into pinned
    // memory slice by slice and enqueue each slice's DMA themselves; the sort kernels, enqueued behind them, do not read the hashes.
    uint8_t* ph = pin_h.as<uint8_t>();
    uint8_t* dh = d_hash.as<uint8_t>();
    const unsigned tc = std::max(1u, std::min(n_threads(n), 8u));
    std::vector<int> rcs(tc, 0);
    std::vector<std::thread> copiers;
    copiers.reserve(tc);
    for (unsigned k = 0; k < tc; ++k) {
        const uint64_t b = n * k / tc, e = n * (k + 1) / tc;
        copiers.emplace_back([=, &rcs] {
            if (e <= b) return;
            cudaSetDevice(ctx->device);
            memcpy(ph + b * 128, reinterpret_cast<const uint8_t*>(hashes) + b * 128, (e - b) * 128);
            if (cudaMemcpyAsync(dh + b * 128, ph + b * 128, (e - b) * 128, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rcs[k] = 1;
        });
    }
    uint64_t skip = 0;
    build_prefix_keys(dur, paths, off, n, pd, p1, p2, &skip);
    const double t1 = now_ms();
    for (auto& t : copiers) t.join();
    for (int r : rcs)
        if (r) {
            ctx->err = "upload of the hash table failed";
            cudaGetLastError();
            return VDF_ERR_CUDA;
        }
    ctx->h2d += n * 128;
    VDF_TRY(gpu_sort(ctx, dur, paths, off, n, pd, p1, p2, skip));
    const double t2 = now_ms();
    if (t_keys) *t_keys = t1 - t0;
    if (t_stage) *t_stage = t2 - t1;
    return VDF_OK;
}

namespace {
// [n][16] u64 rows gathered by a permutation (128-byte rows, 16 bytes per thread)
__global__ void gather_rows_kernel(const uint4* __restrict__ in, const uint32_t* __restrict__ perm, uint64_t n, uint4* __restrict__ out) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;  // 16-byte piece
    if (q >= n * 8) return;
    out[q] = in[(uint64_t)perm[q >> 3] * 8 + (q & 7)];
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

// (tolerance * 1000.0) as u32: truncating, saturating, NaN -> 0 (search_algorithm.rs:82)
uint32_t vdf::tolerance_to_int(double tolerance) {
    const double v = tolerance * 1000.0;
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

namespace {
// keys (ref << 32 | sorted candidate position), sorted -> CSR over the caller's candidate indices
__global__ void ref_csr_kernel(const uint64_t* __restrict__ keys, uint64_t ne, const uint32_t* __restrict__ order, uint64_t n_ref,
                               uint64_t* __restrict__ row_ptr, uint64_t* __restrict__ col_idx) {
    const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= ne) return;
    const uint64_t key = keys[k];
    const uint32_t r = (uint32_t)(key >> 32);
    col_idx[k] = order ? order[(uint32_t)key] : (uint32_t)key;
    // row_ptr[q] = first key of a row >= q: rows (previous key's row, r] start here
    const uint32_t prev = k ? (uint32_t)(keys[k - 1] >> 32) + 1 : 0;
    for (uint64_t q = prev; q <= r; ++q) row_ptr[q] = k;
    if (k + 1 == ne)
        for (uint64_t q = (uint64_t)r + 1; q <= n_ref; ++q) row_ptr[q] = ne;
}
}  // namespace

// out->row_ptr (calloc'ed, n_ref + 1) is filled, out->col_idx allocated: the CSR is built on the device (keys are sorted by
// (ref, sorted candidate position)) with the candidates as the caller's indices
int vdf::ref_keys_to_csr(vdf_ctx* ctx, const uint64_t* d_keys, uint64_t ne, const uint32_t* d_order, uint64_t n_ref, vdf_csr* out) {
    out->col_idx = (uint64_t*)malloc((ne ? ne : 1) * 8);
    if (!out->col_idx) return VDF_ERR_ALLOC;
    if (!ne) return VDF_OK;
    VDF_ALLOC(ctx, ctx->keys_b.ensure((n_ref + 1 + ne) * 8));
    uint64_t* d_rp = ctx->keys_b.as<uint64_t>();
    uint64_t* d_ci = d_rp + n_ref + 1;
    ref_csr_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(d_keys, ne, d_order, n_ref, d_rp, d_ci);
    VDF_LAUNCHED(ctx);
    VDF_ALLOC(ctx, ctx->h_groups.ensure((n_ref + 1 + ne) * 8));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->h_groups.p, d_rp, (n_ref + 1 + ne) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out->row_ptr, ctx->h_groups.p, (n_ref + 1) * 8);
    memcpy(out->col_idx, ctx->h_groups.as<uint64_t>() + n_ref + 1, ne * 8);
    ctx->d2h += (n_ref + 1 + ne) * 8;
    return VDF_OK;
}

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

int vdf_sort_order_device(vdf_ctx* ctx, const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n,
                          uint32_t* d_order_out, uint32_t* d_dur_sorted_out) {
    VDF_TRY(enter(ctx));
    if ((n && (!durations || !path_off || !d_order_out)) || n >= 0xFFFFFF00ull) return VDF_ERR_INVALID;
    if (n == 0) return VDF_OK;
    const double t0 = now_ms();
    VDF_ALLOC(ctx, ctx->pin_c.ensure(n * 20));
    uint64_t* p1 = ctx->pin_c.as<uint64_t>();
    uint64_t* p2 = p1 + n;
    uint32_t* pd = reinterpret_cast<uint32_t*>(p2 + n);
    uint64_t skip = 0;
    build_prefix_keys(durations, path_blob, path_off, n, pd, p1, p2, &skip);
    const double t1 = now_ms();
    VDF_TRY(gpu_sort(ctx, durations, path_blob, path_off, n, pd, p1, p2, skip));
    VDF_CUDA(ctx, cudaMemcpyAsync(d_order_out, ctx->sk_order.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_dur_sorted_out) VDF_CUDA(ctx, cudaMemcpyAsync(d_dur_sorted_out, ctx->in_dur.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->phase_ms[0] = t1 - t0, ctx->phase_ms[1] = now_ms() - t1, ctx->phase_ms[2] = ctx->phase_ms[3] = 0;
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
    *d_hash_sorted = nullptr, *d_dur_sorted = nullptr;
    ctx->phase_ms[0] = ctx->phase_ms[1] = ctx->phase_ms[2] = ctx->phase_ms[3] = 0;
    if (n == 0) return VDF_OK;
    double t_keys = 0, t_stage = 0;
    VDF_TRY(vdf::stage_and_sort(ctx, hashes, durations, path_blob, path_off, n, ctx->in_hash2, ctx->pin_a, &t_keys, &t_stage));
    const double t0 = now_ms();
    uint64_t* dst_h = d_hash_dst;
    if (!dst_h) {
        VDF_ALLOC(ctx, ctx->in_hash.ensure(n * 128));
        dst_h = ctx->in_hash.as<uint64_t>();
    }
    gather_rows_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, ctx->stream>>>(ctx->in_hash2.as<uint4>(), ctx->sk_order.as<uint32_t>(), n,
                                                                                reinterpret_cast<uint4*>(dst_h));
    VDF_LAUNCHED(ctx);
    if (d_dur_dst) VDF_CUDA(ctx, cudaMemcpyAsync(d_dur_dst, ctx->in_dur.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    *d_hash_sorted = dst_h;
    *d_dur_sorted = d_dur_dst ? d_dur_dst : ctx->in_dur.as<uint32_t>();
    // the permutation, widened on the device, straight into the caller's array
    VDF_ALLOC(ctx, ctx->keys_b.ensure(n * 8));
    widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->sk_order.as<uint32_t>(), n, ctx->keys_b.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    VDF_CUDA(ctx, cudaMemcpyAsync(order_out, ctx->keys_b.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d2h += n * 8;
    ctx->phase_ms[0] = t_keys, ctx->phase_ms[1] = t_stage, ctx->phase_ms[2] = now_ms() - t0;
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
    if (ctx->sub_count > 1) return vdf::mgpu_search(ctx, hashes, durations, path_blob, path_off, n, tolerance, out);
    double t_keys = 0, t_stage = 0;
    uint64_t ne = 0;
    if (n) {
        VDF_TRY(vdf::stage_and_sort(ctx, hashes, durations, path_blob, path_off, n, ctx->in_hash, ctx->pin_a, &t_keys, &t_stage));
        const double t2 = now_ms();
        const uint32_t tol_int = vdf::tolerance_to_int(tolerance);
        // Search::from: the table, gathered into sorted order while it is packed
        VDF_TRY(vdf::table_prepare(ctx, ctx->tmp_self, ctx->in_hash.as<uint64_t>(), ctx->sk_order.as<uint32_t>(), ctx->in_dur.as<uint32_t>(), n,
                                   true, false));
        VDF_TRY(with_growing_keys(
            ctx, [&](uint64_t* k, uint64_t cap, uint64_t* cnt) { return vdf::table_search_self(ctx, ctx->tmp_self, tol_int, k, cap, cnt); }, &ne));
        // sorted positions -> the caller's indices (what `entries[i].value.src_path()` resolves to in the reference): on the
        // device, as the group CSR is written
        VDF_TRY(vdf::group_device(ctx, n, ctx->keys_a.as<uint64_t>(), ne, ctx->sk_order.as<uint32_t>(), out));
        ctx->phase_ms[2] = now_ms() - t2;
    } else {
        VDF_TRY(vdf::group_device(ctx, 0, nullptr, 0, nullptr, out));
        ctx->phase_ms[2] = 0;
    }
    ctx->phase_ms[0] = t_keys, ctx->phase_ms[1] = t_stage, ctx->phase_ms[3] = 0;
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
    if (ctx->sub_count > 1)
        return vdf::mgpu_search_refs(ctx, ref_hashes, ref_durations, n_ref, cand_hashes, cand_durations, cand_path_blob, cand_path_off, n_cand,
                                     tolerance, out);
    out->n_rows = n_ref;
    out->row_ptr = (uint64_t*)calloc(n_ref + 1, 8);
    out->col_idx = nullptr;
    if (!out->row_ptr) return VDF_ERR_ALLOC;
    double t_keys = 0, t_stage = 0;
    uint64_t ne = 0;
    const double t2 = now_ms();
    if (n_cand && n_ref) {
        VDF_TRY(vdf::stage_and_sort(ctx, cand_hashes, cand_durations, cand_path_blob, cand_path_off, n_cand, ctx->in_hash, ctx->pin_a, &t_keys, &t_stage));
        VDF_ALLOC(ctx, ctx->in_hash2.ensure(n_ref * 128));
        VDF_ALLOC(ctx, ctx->in_dur2.ensure(n_ref * 4));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_hash2.p, ref_hashes, n_ref * 128, cudaMemcpyHostToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_dur2.p, ref_durations, n_ref * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += n_ref * 132;
        const uint32_t tol_int = vdf::tolerance_to_int(tolerance);
        VDF_TRY(vdf::table_prepare(ctx, ctx->tmp_cand, ctx->in_hash.as<uint64_t>(), ctx->sk_order.as<uint32_t>(), ctx->in_dur.as<uint32_t>(),
                                   n_cand, false, true));
        VDF_TRY(with_growing_keys(
            ctx,
            [&](uint64_t* k, uint64_t cap, uint64_t* cnt) {
                return vdf::table_search_refs(ctx, ctx->tmp_cand, 0, ctx->in_hash2.as<uint64_t>(), ctx->in_dur2.as<uint32_t>(), n_ref, tol_int, k, cap,
                                              cnt);
            },
            &ne));
    }
    VDF_TRY(vdf::ref_keys_to_csr(ctx, ctx->keys_a.as<uint64_t>(), ne, ctx->sk_order.as<uint32_t>(), n_ref, out));
    ctx->phase_ms[0] = t_keys, ctx->phase_ms[1] = t_stage, ctx->phase_ms[2] = now_ms() - t2 - t_keys - t_stage, ctx->phase_ms[3] = 0;
    return VDF_OK;
}

int vdf_ctx_last_phases(const vdf_ctx* ctx, double* ms4) {
    if (!ctx || !ms4) return VDF_ERR_INVALID;
    for (int k = 0; k < 4; ++k) ms4[k] = ctx->phase_ms[k];
    return VDF_OK;
}

}  // extern "C"
