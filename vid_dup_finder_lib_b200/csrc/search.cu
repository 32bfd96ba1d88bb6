// search.cu -- tiled XOR+POPC Hamming search.
//
// Replaces the comparison loops of Search::search_self (search_algorithm.rs:81-117,140-156) and
// Search::search_one (search_algorithm.rs:63-77) with VideoHash::hamming_distance (video_hash.rs:311-317).
//
// Data layout in HBM: the sorted hash table is re-tiled once per call into
//     tiles[tile][word32 0..31][hash 0..127]          (16 KB per tile, zero padded)
// so that (a) one tile is ONE contiguous 16 KB block that a single cp.async.bulk (TMA bulk copy, SASS UBLKCP)
// lands in shared memory, completing on an mbarrier, and (b) a warp reading word w of 4 consecutive hashes per
// lane does conflict-free 128-bit shared loads.  A CTA owns a row tile (128 hashes) and streams a chunk of column
// tiles through a 3-stage shared-memory ring; each of its 256 threads keeps an 8x8 block of pair accumulators
// in registers.  Pairs under the tolerance are appended (atomic counter) to an edge buffer as u64 keys
// (row << 32 | col); the window test lo[row] <= col < hi[row] is evaluated only for those rare pairs.
// The per-row windows make the same kernel serve search_self (lo = i+1, hi = end of the 1.1x duration window)
// and search_with_references (lo/hi = the 0.95x..1.05x duration slice of the candidate table).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace vdf {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA bulk copy global -> shared (1-D, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ re-tiling
// in: [n][32] u32 (hash-major, 128 B per hash); out: [ceil(n/128)][32][128] u32.  perm (optional) gathers rows.
__global__ void __launch_bounds__(256) retile_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm,
                                                     uint64_t n, uint32_t* __restrict__ out) {
    __shared__ uint32_t s[kWords32][kTile + 1];
    const uint64_t tile = blockIdx.x;
    const int tid = threadIdx.x;
    // 128 hashes x 8 uint4 = 1024 uint4 loads, coalesced along the hash
    for (int q = tid; q < kTile * 8; q += 256) {
        int h = q >> 3, c = q & 7;
        uint64_t g = tile * kTile + h;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (g < n) {
            uint64_t src = perm ? perm[g] : g;
            v = reinterpret_cast<const uint4*>(in)[src * 8 + c];
        }
        s[4 * c + 0][h] = v.x;
        s[4 * c + 1][h] = v.y;
        s[4 * c + 2][h] = v.z;
        s[4 * c + 3][h] = v.w;
    }
    __syncthreads();
    uint32_t* o = out + tile * kTileWords;
    for (int q = tid; q < kTileWords; q += 256) o[q] = s[q >> 7][q & 127];
}

// ------------------------------------------------------------------------------------------------ windows
// Rust `f64 as u32` == cvt.rzi.u32.f64 (saturating, NaN -> 0)
__device__ __forceinline__ uint32_t f64_as_u32(double v) { return __double2uint_rz(v); }

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* d, uint32_t n, uint32_t v) {  // first k: d[k] >= v
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (d[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* d, uint32_t n, uint32_t v) {  // first k: d[k] > v
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (d[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// search_self: candidate j of target i must satisfy i < j and dur[j] <= (f64(dur[i]) * 1.1) as u32
// (search_algorithm.rs:99,110).  n_pad = rows rounded up to whole tiles; padded rows get an empty window.
__global__ void self_window_kernel(const uint32_t* __restrict__ dur, uint32_t n, uint32_t n_pad,
                                   uint32_t* __restrict__ row_lo, uint32_t* __restrict__ row_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i >= n) {
        row_lo[i] = 0;
        row_hi[i] = 0;
        return;
    }
    uint32_t thresh = f64_as_u32((double)dur[i] * 1.1);
    row_lo[i] = i + 1;
    row_hi[i] = upper_bound_u32(dur, n, thresh);
}

// search_with_references: duration_slice (search_algorithm.rs:173-185), rows in the (sorted-by-lo) internal order
__global__ void ref_window_kernel(const uint32_t* __restrict__ cand_dur, uint32_t n_cand,
                                  const uint32_t* __restrict__ ref_dur, const uint32_t* __restrict__ perm, uint32_t n_ref,
                                  uint32_t n_pad, uint32_t* __restrict__ row_lo, uint32_t* __restrict__ row_hi) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    if (r >= n_ref) {
        row_lo[r] = 0;
        row_hi[r] = 0;
        return;
    }
    uint32_t d = ref_dur[perm ? perm[r] : r];
    uint32_t lo_d = f64_as_u32((double)d * 0.95), hi_d = f64_as_u32((double)d * 1.05);
    row_lo[r] = lower_bound_u32(cand_dur, n_cand, lo_d);
    row_hi[r] = upper_bound_u32(cand_dur, n_cand, hi_d);
}

// sort key for the references: start of their duration slice (so that a row tile's slices overlap)
__global__ void ref_sortkey_kernel(const uint32_t* __restrict__ ref_dur, uint32_t n_ref, uint32_t* __restrict__ key,
                                   uint32_t* __restrict__ idx) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ref) return;
    key[r] = ref_dur[r];
    idx[r] = r;
}

// per row tile: the range of column tiles any of its rows can match; misc[0] = max span, misc[1..2] = pair count
__global__ void tile_range_kernel(const uint32_t* __restrict__ row_lo, const uint32_t* __restrict__ row_hi,
                                  uint32_t n_row_tiles, uint2* __restrict__ tile_range,
                                  unsigned long long* __restrict__ misc) {
    uint32_t I = blockIdx.x;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    unsigned long long pairs = 0;
    for (int r = threadIdx.x; r < kTile; r += blockDim.x) {
        uint32_t a = row_lo[I * kTile + r], b = row_hi[I * kTile + r];
        if (a < b) {
            lo = min(lo, a);
            hi = max(hi, b);
            pairs += b - a;
        }
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    }
    __shared__ uint32_t slo[4], shi[4];
    __shared__ unsigned long long sp[4];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) slo[w] = lo, shi[w] = hi, sp[w] = pairs;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) lo = min(lo, slo[k]), hi = max(hi, shi[k]), pairs += sp[k];
        uint2 r = make_uint2(0, 0);
        if (lo < hi) r = make_uint2(lo / kTile, (hi + kTile - 1) / kTile);
        tile_range[I] = r;
        atomicMax(&misc[0], (unsigned long long)(r.y - r.x));
        atomicAdd(&misc[1], pairs);
    }
}

// ------------------------------------------------------------------------------------------------ the hot kernel
constexpr int kStages = 3;
constexpr int kHamThreads = 256;
constexpr size_t kHamSmem = (size_t)(1 + kStages) * kTileWords * 4 + 64;

struct HamParams {
    const uint32_t* row_tiles;
    const uint32_t* col_tiles;
    const uint32_t* row_lo;
    const uint32_t* row_hi;
    const uint32_t* row_id;  // nullable: output row index = internal row index
    const uint2* tile_range;
    uint64_t* keys;
    unsigned long long* counter;
    uint64_t capacity;
    uint64_t col_base;
    uint32_t chunk;  // column tiles per CTA
    uint32_t tol;
    uint32_t rank, world;
    uint32_t one;  // = 1, opaque to the compiler: IMAD acc = x * one + acc keeps the accumulate adds on the FMA pipe
};

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// integer multiply-add: issues on the FMA pipe, leaving the ALU pipe to the LOP3s
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// VARIANT 0: acc += popc(a ^ b) per 32-bit word: 32 POPC per pair (the reference's 16 x POPCNT64).
// VARIANT 1: a carry-save adder folds two XOR words into a running "ones" word and pops only the carries:
//            17 POPC + 64 LOP3 per pair -- trades quarter-rate POPC issue slots for full-rate LOP3.
template <int VARIANT>
__global__ void __launch_bounds__(kHamThreads, VARIANT == 0 ? 2 : 1) hamming_tiles_kernel(const HamParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* sA = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sB = sA + kTileWords;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStages * kTileWords);  // [kStages] col stages, [kStages] = row tile

    const uint32_t I = blockIdx.x, c = blockIdx.y;  // x = row tile (up to 2^31-1), y = chunk
    if (p.world > 1 && ((I + c) % p.world) != p.rank) return;  // block-cyclic deal of (row tile, chunk) units to ranks
    const uint2 rng = p.tile_range[I];
    const uint32_t jt0 = rng.x + c * p.chunk;
    if (jt0 >= rng.y) return;
    const uint32_t ntiles = min(p.chunk, rng.y - jt0);
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s <= kStages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bars[kStages], kTileWords * 4);
        bulk_g2s(sA, p.row_tiles + (size_t)I * kTileWords, kTileWords * 4, &bars[kStages]);
        for (uint32_t s = 0; s < kStages && s < ntiles; ++s) {
            mbar_expect_tx(&bars[s], kTileWords * 4);
            bulk_g2s(sB + s * kTileWords, p.col_tiles + (size_t)(jt0 + s) * kTileWords, kTileWords * 4, &bars[s]);
        }
    }
    // thread (ty, tx) owns rows {4ty..4ty+3, 64+4ty..} x cols {4tx..4tx+3, 64+4tx..}
    const int ty = tid >> 4, tx = tid & 15;
    const uint32_t* pa = sA + ty * 4;
    mbar_wait(&bars[kStages], 0);

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t s = t % kStages;
        mbar_wait(&bars[s], (t / kStages) & 1);
        const uint32_t* pb = sB + s * kTileWords + tx * 4;

        uint32_t acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0;

        if (VARIANT == 0) {
#pragma unroll 2
            for (int w = 0; w < kWords32; ++w) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(pa + w * kTile);
                const uint4 a1 = *reinterpret_cast<const uint4*>(pa + w * kTile + 64);
                const uint4 b0 = *reinterpret_cast<const uint4*>(pb + w * kTile);
                const uint4 b1 = *reinterpret_cast<const uint4*>(pb + w * kTile + 64);
                const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] += __popc(a[i] ^ b[j]);
            }
        } else {
            uint32_t ones[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) ones[i][j] = 0;
#pragma unroll 1
            for (int w = 0; w < kWords32; w += 2) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(pa + w * kTile);
                const uint4 a1 = *reinterpret_cast<const uint4*>(pa + w * kTile + 64);
                const uint4 b0 = *reinterpret_cast<const uint4*>(pb + w * kTile);
                const uint4 b1 = *reinterpret_cast<const uint4*>(pb + w * kTile + 64);
                const uint4 c0 = *reinterpret_cast<const uint4*>(pa + (w + 1) * kTile);
                const uint4 c1 = *reinterpret_cast<const uint4*>(pa + (w + 1) * kTile + 64);
                const uint4 d0 = *reinterpret_cast<const uint4*>(pb + (w + 1) * kTile);
                const uint4 d1 = *reinterpret_cast<const uint4*>(pb + (w + 1) * kTile + 64);
                const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const uint32_t cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const uint32_t d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t x0 = a[i] ^ b[j], x1 = cc[i] ^ d[j];
                        const uint32_t carry = maj3(ones[i][j], x0, x1);
                        ones[i][j] = xor3(ones[i][j], x0, x1);
                        acc[i][j] = imad(__popc(carry), p.one, acc[i][j]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = imad(acc[i][j], p.one + p.one, __popc(ones[i][j]));
        }

        // rare path: pairs under the tolerance -> window test -> append
        // (one min-reduction and a single branch keep the hot path's instruction footprint small)
        uint32_t best = acc[0][0];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) best = min(best, acc[i][j]);
        if (best <= p.tol) {
            uint64_t mask = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) mask |= (uint64_t)(acc[i][j] <= p.tol) << (i * 8 + j);
            const uint32_t col0 = (jt0 + t) * kTile;
            while (mask) {
                const int b = __ffsll((long long)mask) - 1;
                mask &= mask - 1;
                const int i = b >> 3, j = b & 7;
                const uint32_t gi = I * kTile + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
                const uint32_t gj = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (gj >= p.row_lo[gi] && gj < p.row_hi[gi]) {
                    const unsigned long long slot = atomicAdd(p.counter, 1ull);
                    if (slot < p.capacity) {
                        const uint64_t rid = p.row_id ? p.row_id[gi] : gi;
                        p.keys[slot] = (rid << 32) | (uint64_t)(gj + p.col_base);
                    }
                }
            }
        }

        __syncthreads();  // every thread is done with stage s
        if (tid == 0 && t + kStages < ntiles) {
            mbar_expect_tx(&bars[s], kTileWords * 4);
            bulk_g2s(sB + s * kTileWords, p.col_tiles + (size_t)(jt0 + t + kStages) * kTileWords, kTileWords * 4, &bars[s]);
        }
    }
}

// VARIANT 2: the carry-save scheme of variant 1 on an 8x4 register block (two passes over the halves of each
// column tile): ~half the registers, so two CTAs (16 warps) share an SM and hide the LOP3 -> POPC -> IMAD latencies.
__global__ void __launch_bounds__(kHamThreads, 2) hamming_tiles_csa4_kernel(const HamParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint32_t* sA = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sB = sA + kTileWords;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStages * kTileWords);

    const uint32_t I = blockIdx.x, c = blockIdx.y;
    if (p.world > 1 && ((I + c) % p.world) != p.rank) return;
    const uint2 rng = p.tile_range[I];
    const uint32_t jt0 = rng.x + c * p.chunk;
    if (jt0 >= rng.y) return;
    const uint32_t ntiles = min(p.chunk, rng.y - jt0);
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int s = 0; s <= kStages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bars[kStages], kTileWords * 4);
        bulk_g2s(sA, p.row_tiles + (size_t)I * kTileWords, kTileWords * 4, &bars[kStages]);
        for (uint32_t s = 0; s < kStages && s < ntiles; ++s) {
            mbar_expect_tx(&bars[s], kTileWords * 4);
            bulk_g2s(sB + s * kTileWords, p.col_tiles + (size_t)(jt0 + s) * kTileWords, kTileWords * 4, &bars[s]);
        }
    }
    // thread (ty, tx): rows {4ty..4ty+3, 64+4ty..}; per pass h: cols 64h + 4tx..4tx+3
    const int ty = tid >> 4, tx = tid & 15;
    const uint32_t* pa = sA + ty * 4;
    mbar_wait(&bars[kStages], 0);

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t s = t % kStages;
        mbar_wait(&bars[s], (t / kStages) & 1);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const uint32_t* pb = sB + s * kTileWords + h * 64 + tx * 4;
            uint32_t acc[8][4], ones[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0, ones[i][j] = 0;
#pragma unroll 1
            for (int w = 0; w < kWords32; w += 2) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(pa + w * kTile);
                const uint4 a1 = *reinterpret_cast<const uint4*>(pa + w * kTile + 64);
                const uint4 b0 = *reinterpret_cast<const uint4*>(pb + w * kTile);
                const uint4 c0 = *reinterpret_cast<const uint4*>(pa + (w + 1) * kTile);
                const uint4 c1 = *reinterpret_cast<const uint4*>(pa + (w + 1) * kTile + 64);
                const uint4 d0 = *reinterpret_cast<const uint4*>(pb + (w + 1) * kTile);
                const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const uint32_t b[4] = {b0.x, b0.y, b0.z, b0.w};
                const uint32_t cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const uint32_t d[4] = {d0.x, d0.y, d0.z, d0.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t x0 = a[i] ^ b[j], x1 = cc[i] ^ d[j];
                        const uint32_t carry = maj3(ones[i][j], x0, x1);
                        ones[i][j] = xor3(ones[i][j], x0, x1);
                        acc[i][j] = imad(__popc(carry), p.one, acc[i][j]);
                    }
            }
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = imad(acc[i][j], p.one + p.one, __popc(ones[i][j]));
                    best = min(best, acc[i][j]);
                }
            if (best <= p.tol) {
                uint32_t mask = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) mask |= (uint32_t)(acc[i][j] <= p.tol) << (i * 4 + j);
                const uint32_t col0 = (jt0 + t) * kTile + h * 64 + tx * 4;
                while (mask) {
                    const int b = __ffs((int)mask) - 1;
                    mask &= mask - 1;
                    const int i = b >> 2, j = b & 3;
                    const uint32_t gi = I * kTile + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
                    const uint32_t gj = col0 + j;
                    if (gj >= p.row_lo[gi] && gj < p.row_hi[gi]) {
                        const unsigned long long slot = atomicAdd(p.counter, 1ull);
                        if (slot < p.capacity) {
                            const uint64_t rid = p.row_id ? p.row_id[gi] : gi;
                            p.keys[slot] = (rid << 32) | (uint64_t)(gj + p.col_base);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && t + kStages < ntiles) {
            mbar_expect_tx(&bars[s], kTileWords * 4);
            bulk_g2s(sB + s * kTileWords, p.col_tiles + (size_t)(jt0 + t + kStages) * kTileWords, kTileWords * 4, &bars[s]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ peer exchange
// End of an exchange call: publish this rank's match count into every rank's buffer and signal; then wait until all ranks
// have signalled.  The key stores of the pair kernel are ordered before the signal by the kernel boundary and a system-scope
// fence; the wait is an acquire load at system scope.  `arrived` only ever grows (world per call on this half), so nothing
// is reset across ranks.  A wait longer than `timeout_ns` (a dead peer) sets *timed_out instead of hanging the GPU.
__global__ void peer_barrier_kernel(PeerPtrs pp, const unsigned long long* __restrict__ local_count, unsigned long long target,
                                    unsigned long long timeout_ns, uint32_t* timed_out) {
    if (threadIdx.x < pp.world) {
        *pp.seg_count(threadIdx.x, pp.rank) = *local_count;
        __threadfence_system();
        atomicAdd_system(pp.arrived(threadIdx.x), 1ull);
    }
    if (threadIdx.x == 0) {
        unsigned long long t0, t1, seen;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const unsigned long long* mine = pp.arrived(pp.rank);
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= target) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                *timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
        __threadfence_system();
    }
}
// the world segments of this rank's buffer -> one contiguous list; out[0] = total, out[1] = largest segment count
__global__ void peer_compact_kernel(PeerPtrs pp, uint64_t* __restrict__ dst, uint64_t dst_cap, unsigned long long* __restrict__ out) {
    unsigned long long off[kMaxPeers + 1];
    unsigned long long biggest = 0;
    off[0] = 0;
    for (uint32_t w = 0; w < pp.world; ++w) {
        const unsigned long long c = *pp.seg_count(pp.rank, w);
        biggest = c > biggest ? c : biggest;
        off[w + 1] = off[w] + (c < pp.seg_cap ? c : pp.seg_cap);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = off[pp.world], out[1] = biggest;
    const unsigned long long total = off[pp.world] < dst_cap ? off[pp.world] : dst_cap;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < total;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t w = 0;
        while (k >= off[w + 1]) ++w;
        dst[k] = pp.keys(pp.rank, w)[k - off[w]];
    }
}

// ------------------------------------------------------------------------------------------------ host side
// bits needed for values < n
int bits_for(uint64_t n) {
    int b = 1;
    while (b < 64 && (1ull << b) < n) ++b;
    return b;
}

// ascending radix sort of the bits [0, end_bit) of the keys: every key of the search is (index < n) << 32 | index, so the top
// 32 - bits(n) bits are zero and one or two of the eight 8-bit passes can be left out
int sort_keys(vdf_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, int end_bit) {
    if (n == 0) return VDF_OK;
    size_t tmp = 0;
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp, d_in, d_out, (size_t)n, 0, end_bit, ctx->stream));
    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
    VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp, d_in, d_out, (size_t)n, 0, end_bit, ctx->stream));
    ctx->launches += 4;  // cub's histogram + onesweep passes (approximate; library kernels)
    return VDF_OK;
}

static int set_ham_attrs(vdf_ctx* ctx) {  // cudaFuncSetAttribute is per device: once per context
    if (ctx->ham_attrs) return VDF_OK;
    VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tiles_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHamSmem));
    VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tiles_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHamSmem));
    VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tiles_csa4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHamSmem));
    ctx->ham_attrs = true;
    return VDF_OK;
}

// one side of the pair matrix in the form the selected kernel reads (column role differs from row role for variant 6 only)
static int pack_side(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, bool column_role, Packed& out,
                     uint32_t* d_pads_or) {
    const int v = ctx->search_variant;
    if (v == 6) return tc6_pack(ctx, d_hash, perm, n, column_role, out, d_pads_or);
    if (v == 5) return tc5_pack(ctx, d_hash, perm, n, out);
    const uint32_t T = (uint32_t)((n + kTile - 1) / kTile);
    VDF_ALLOC(ctx, out.tiles.ensure((size_t)T * kTileWords * 4));
    retile_kernel<<<T, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_hash), perm, n, out.tiles.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    out.n = n, out.variant = v, out.as_columns = false;
    return VDF_OK;
}

// `Search::from(hashes)` (search_algorithm.rs:188-198): the sorted table in the kernels' layout.  Enqueue only.
int table_prepare(vdf_ctx* ctx, Table& t, const uint64_t* d_hash, const uint32_t* d_perm, const uint32_t* d_dur_sorted, uint64_t n,
                  bool for_self, bool for_cand) {
    t.n = n, t.d_hash = d_hash, t.d_perm = d_perm, t.d_dur = d_dur_sorted;
    t.self_plan.valid = false;
    t.pads = 0, t.pads_known = false;
    t.rows.variant = t.cols.variant = -1;
    if (n == 0) return VDF_OK;
    VDF_ALLOC(ctx, t.meta.ensure(64));
    VDF_CUDA(ctx, cudaMemsetAsync(t.meta.p, 0, 64, ctx->stream));
    const bool v6 = ctx->search_variant == 6;
    if (for_self || !v6) VDF_TRY(pack_side(ctx, d_hash, d_perm, n, false, t.rows, t.meta.as<uint32_t>()));
    if (v6 && (for_self || for_cand)) VDF_TRY(pack_side(ctx, d_hash, d_perm, n, true, t.cols, t.meta.as<uint32_t>()));
    return VDF_OK;
}

static bool plan_matches(const vdf_ctx* ctx, const Plan& pl) {
    return pl.valid && pl.variant == ctx->search_variant && pl.rank == ctx->rank && pl.world == ctx->world && pl.tc_chunk == ctx->tc_chunk &&
           pl.unit_order == ctx->tc_unit_order && pl.tc_fold == ctx->tc_fold;
}

// windows are in pl.row_lo / row_hi: tile ranges, statistics, work units -- then ONE read-back of everything the host needs
// to size the launch: max span, pairs in the windows, units, pad bits of the operands
static int plan_finish(vdf_ctx* ctx, Plan& pl, uint32_t n_row_tiles, uint32_t n_col_tiles, Table* t_a, uint32_t extra_pads_slot) {
    pl.n_row_tiles = n_row_tiles, pl.n_col_tiles = n_col_tiles;
    pl.variant = ctx->search_variant, pl.rank = ctx->rank, pl.world = ctx->world, pl.tc_chunk = ctx->tc_chunk;
    pl.unit_order = ctx->tc_unit_order, pl.tc_fold = ctx->tc_fold;
    VDF_ALLOC(ctx, pl.tile_range.ensure((size_t)(n_row_tiles + 1) * sizeof(uint2)));
    unsigned long long* stats = pl.stats.as<unsigned long long>();
    tile_range_kernel<<<n_row_tiles, 128, 0, ctx->stream>>>(pl.row_lo.as<uint32_t>(), pl.row_hi.as<uint32_t>(), n_row_tiles,
                                                             pl.tile_range.as<uint2>(), stats);
    VDF_LAUNCHED(ctx);
    if (pl.variant >= 5) VDF_TRY(tc_plan_units(ctx, pl));
    VDF_ALLOC(ctx, ctx->h_misc.ensure(256));
    unsigned long long* h = ctx->h_misc.as<unsigned long long>();
    h[8] = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(h, stats, 64, cudaMemcpyDeviceToHost, ctx->stream));
    if (t_a && !t_a->pads_known) VDF_CUDA(ctx, cudaMemcpyAsync(h + 8, t_a->meta.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pl.max_span = (uint32_t)h[0], pl.pairs = h[1], pl.n_units = (uint32_t)h[2];
    if (t_a && !t_a->pads_known) t_a->pads = (uint32_t)h[8], t_a->pads_known = true;
    uint32_t pads = t_a ? t_a->pads : 0;
    if (extra_pads_slot) pads |= (uint32_t)h[extra_pads_slot];
    pl.fold = pl.variant == 6 && ctx->tc_fold != 0 && pads == 0;
    if (pl.variant >= 5 && pl.n_units > 0x3FFFFFFFu) {
        ctx->err = "too many work units for one launch";
        return VDF_ERR_INVALID;
    }
    pl.valid = true;
    return VDF_OK;
}

static int plan_alloc(vdf_ctx* ctx, Plan& pl, uint32_t rows_padded) {
    VDF_ALLOC(ctx, pl.row_lo.ensure((size_t)rows_padded * 4));
    VDF_ALLOC(ctx, pl.row_hi.ensure((size_t)rows_padded * 4));
    VDF_ALLOC(ctx, pl.stats.ensure(64));
    VDF_CUDA(ctx, cudaMemsetAsync(pl.stats.p, 0, 64, ctx->stream));
    pl.valid = false;
    return VDF_OK;
}

// launch the pair kernel of the plan, close the exchange if one is on, read the match count (the one host round trip of a
// search on a prepared table), sort the keys
static int run_plan(vdf_ctx* ctx, const Plan& pl, const Packed& rows, const Packed& cols, const uint32_t* row_id, uint64_t col_base,
                    uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    *n_out = 0;
    VDF_ALLOC(ctx, ctx->misc.ensure(64));
    VDF_ALLOC(ctx, ctx->raw_keys.ensure((size_t)(capacity ? capacity : 1) * 8));
    unsigned long long* misc = ctx->misc.as<unsigned long long>();  // [2] match counter, [3] time-out flag, [4..5] exchange totals
    VDF_CUDA(ctx, cudaMemsetAsync(misc, 0, 64, ctx->stream));
    PeerPtrs pp;
    pp.world = 0;
    if (ctx->exchange) {  // matches of ALL ranks arrive in this rank's peer buffer (common.cuh: PeerExchange)
        if (ctx->search_variant != 6) {
            ctx->err = "peer exchange: only the default kernel (search_variant 6) stores to peer buffers";
            return VDF_ERR_INVALID;
        }
        if (!ctx->peer.world) {
            ctx->err = "peer exchange: call vdf_peer_alloc / vdf_peer_open first";
            return VDF_ERR_INVALID;
        }
        if (ctx->peer_dead) {
            ctx->err = "peer exchange: an earlier exchange failed on this buffer; vdf_peer_close, then allocate and open again on every rank";
            return VDF_ERR_INVALID;
        }
        pp = ctx->peer.ptrs((uint32_t)(ctx->peer.epoch & 1));
    }
    const bool any_work = pl.variant >= 5 ? pl.n_units != 0 : pl.max_span != 0;
    if (!any_work && !pp.world) return VDF_OK;
    // from here on a failure leaves the ranks' exchange counters out of step
    struct DeadGuard {
        vdf_ctx* c;
        bool armed;
        ~DeadGuard() {
            if (armed) c->peer_dead = true;
        }
    } guard{ctx, pp.world != 0};
    if (any_work) {
        if (pl.variant >= 5) {
            VDF_TRY(tc_launch(ctx, pl, rows, cols, row_id, col_base, tol, capacity, misc + 2));
        } else {
            VDF_TRY(set_ham_attrs(ctx));
            // chunk: enough column tiles per CTA to amortise the row-tile load, small enough to balance 148 SMs
            uint32_t chunk = 32;
            while (chunk > 4 && (uint64_t)pl.n_row_tiles * ((pl.max_span + chunk - 1) / chunk) < (uint64_t)ctx->sm_count * 8 * pl.world) chunk >>= 1;
            HamParams p;
            p.row_tiles = rows.tiles.as<uint32_t>();
            p.col_tiles = cols.tiles.as<uint32_t>();
            p.row_lo = pl.row_lo.as<uint32_t>();
            p.row_hi = pl.row_hi.as<uint32_t>();
            p.row_id = row_id;
            p.tile_range = pl.tile_range.as<uint2>();
            p.keys = ctx->raw_keys.as<uint64_t>();
            p.counter = misc + 2;
            p.capacity = capacity;
            p.col_base = col_base;
            p.chunk = chunk;
            p.tol = tol;
            p.rank = pl.rank;
            p.world = pl.world;
            p.one = 1;
            dim3 grid(pl.n_row_tiles, (pl.max_span + chunk - 1) / chunk);
            while (grid.y > 65535) {  // only reachable beyond ~268M columns
                chunk *= 2;
                p.chunk = chunk;
                grid.y = (pl.max_span + chunk - 1) / chunk;
            }
            kt_begin(ctx, 0);
            if (pl.variant == 2) hamming_tiles_csa4_kernel<<<grid, kHamThreads, kHamSmem, ctx->stream>>>(p);
            else if (pl.variant == 1) hamming_tiles_kernel<1><<<grid, kHamThreads, kHamSmem, ctx->stream>>>(p);
            else hamming_tiles_kernel<0><<<grid, kHamThreads, kHamSmem, ctx->stream>>>(p);
            kt_end(ctx, 0);
            VDF_LAUNCHED(ctx);
        }
    }
    VDF_ALLOC(ctx, ctx->h_misc.ensure(256));
    unsigned long long* h = ctx->h_misc.as<unsigned long long>();
    const uint64_t* raw = ctx->raw_keys.as<uint64_t>();
    unsigned long long cnt = 0;
    if (pp.world) {
        const unsigned long long target = (unsigned long long)pp.world * (ctx->peer.epoch / 2 + 1);
        // a dead peer must not hang the GPU; a slow one must not be mistaken for dead: the allowance grows with this rank's share
        const unsigned long long timeout_ms = ctx->peer_timeout_ms ? ctx->peer_timeout_ms : 20000ull + (pl.pairs >> 26);
        peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(pp, misc + 2, target, timeout_ms * 1000000ull, reinterpret_cast<uint32_t*>(misc + 3));
        VDF_LAUNCHED(ctx);
        ctx->peer.epoch += 1;
        const uint64_t raw_cap = capacity ? capacity : 1;
        peer_compact_kernel<<<64, 256, 0, ctx->stream>>>(pp, ctx->raw_keys.as<uint64_t>(), raw_cap, misc + 4);
        VDF_LAUNCHED(ctx);
        VDF_CUDA(ctx, cudaMemcpyAsync(h, misc, 64, cudaMemcpyDeviceToHost, ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if ((uint32_t)h[3]) {
            ctx->err = "peer exchange: a rank did not finish its search within " + std::to_string(timeout_ms) + " ms (option peer_timeout_ms)";
            return VDF_ERR_CUDA;
        }
        guard.armed = false;  // every rank passed the barrier: the counters agree again, whatever the counts say
        cnt = h[4];
        *n_out = cnt;
        if (h[5] > pp.seg_cap) {  // some rank found more than a segment holds: every rank sees this alike
            *n_out = h[5] * pp.world;
            ctx->err = "edge buffer overflow: " + std::to_string(h[5]) + " matches on one rank > segment capacity " + std::to_string(pp.seg_cap);
            return VDF_ERR_EDGE_OVERFLOW;
        }
    } else {
        VDF_CUDA(ctx, cudaMemcpyAsync(h, misc, 64, cudaMemcpyDeviceToHost, ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cnt = h[2];
        *n_out = cnt;
    }
    if (cnt > capacity) {
        ctx->err = "edge buffer overflow: " + std::to_string(cnt) + " matches > capacity " + std::to_string(capacity);
        return VDF_ERR_EDGE_OVERFLOW;
    }
    return sort_keys(ctx, raw, d_keys_out, cnt, 32 + bits_for(rows.n));  // keys: (row index < rows.n) << 32 | column index
}

// `Search::search_self` (search_algorithm.rs:81-171, comparison part) on a prepared table
int table_search_self(vdf_ctx* ctx, Table& t, uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    *n_out = 0;
    const uint64_t n = t.n;
    if (n == 0) return VDF_OK;  // search_algorithm.rs:88-90
    if (n >= 0xFFFFFF00ull) {
        ctx->err = "n must be < 2^32";
        return VDF_ERR_INVALID;
    }
    const int v = ctx->search_variant;
    if (t.rows.variant != v || (v == 6 && t.cols.variant != v)) {  // the kernel choice changed since the table was packed
        VDF_TRY(table_prepare(ctx, t, t.d_hash, t.d_perm, t.d_dur, n, true, true));
    }
    const uint32_t T = (uint32_t)((n + kTile - 1) / kTile);
    Plan& pl = t.self_plan;
    if (!plan_matches(ctx, pl)) {
        const uint32_t n_pad = T * kTile;
        VDF_TRY(plan_alloc(ctx, pl, n_pad));
        self_window_kernel<<<(n_pad + 255) / 256, 256, 0, ctx->stream>>>(t.d_dur, (uint32_t)n, n_pad, pl.row_lo.as<uint32_t>(), pl.row_hi.as<uint32_t>());
        VDF_LAUNCHED(ctx);
        VDF_TRY(plan_finish(ctx, pl, T, T, &t, 0));
    }
    return run_plan(ctx, pl, t.rows, v == 6 ? t.cols : t.rows, nullptr, 0, tol, d_keys_out, capacity, n_out);
}

int search_self_device(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* d_dur, uint64_t n, uint32_t tol,
                       uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    *n_out = 0;
    if (n == 0) return VDF_OK;
    if (n >= 0xFFFFFF00ull) {
        ctx->err = "n must be < 2^32";
        return VDF_ERR_INVALID;
    }
    VDF_TRY(table_prepare(ctx, ctx->tmp_self, d_hash, nullptr, d_dur, n, true, false));
    return table_search_self(ctx, ctx->tmp_self, tol, d_keys_out, capacity, n_out);
}

// `Search::search_with_references` (search_algorithm.rs:40-53,63-77,173-185) on a prepared candidate table
int table_search_refs(vdf_ctx* ctx, Table& cand, uint64_t cand_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                      uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    *n_out = 0;
    const uint64_t n_cand = cand.n;
    const bool exchange = ctx->exchange != 0;
    if ((n_cand == 0 || n_ref == 0) && !exchange) return VDF_OK;
    if (n_cand >= 0xFFFFFF00ull || n_ref >= 0xFFFFFF00ull || cand_base + n_cand > 0xFFFFFFFFull) {
        ctx->err = "indices must fit in 32 bits";
        return VDF_ERR_INVALID;
    }
    const int v = ctx->search_variant;
    Packed& cols = v == 6 ? cand.cols : cand.rows;
    if (n_cand && cols.variant != v) VDF_TRY(table_prepare(ctx, cand, cand.d_hash, cand.d_perm, cand.d_dur, n_cand, false, true));
    const uint32_t TC = (uint32_t)((n_cand + kTile - 1) / kTile);
    const uint32_t TR = (uint32_t)((n_ref + kTile - 1) / kTile);
    const uint32_t r_pad = TR * kTile;
    Plan& pl = ctx->ref_plan;
    VDF_TRY(plan_alloc(ctx, pl, r_pad ? r_pad : kTile));
    pl.n_units = 0, pl.max_span = 0, pl.pairs = 0, pl.variant = v;
    uint32_t* perm = nullptr;
    if (n_cand && n_ref) {
        VDF_ALLOC(ctx, ctx->ref_key.ensure((size_t)n_ref * 4 * 3));
        VDF_ALLOC(ctx, ctx->ref_perm.ensure((size_t)r_pad * 4));
        // order the references by duration so that the 128 rows of a tile have overlapping slices
        uint32_t* key_in = ctx->ref_key.as<uint32_t>();
        uint32_t* key_out = key_in + n_ref;
        uint32_t* idx_in = key_out + n_ref;
        perm = ctx->ref_perm.as<uint32_t>();
        ref_sortkey_kernel<<<(uint32_t)((n_ref + 255) / 256), 256, 0, ctx->stream>>>(d_ref_dur, (uint32_t)n_ref, key_in, idx_in);
        VDF_LAUNCHED(ctx);
        size_t tmp = 0;
        VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, key_in, key_out, idx_in, perm, (size_t)n_ref, 0, 32, ctx->stream));
        VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
        VDF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp, key_in, key_out, idx_in, perm, (size_t)n_ref, 0, 32, ctx->stream));
        ctx->launches += 3;
        // pad bits of the references go to stats[3] (read back with the plan)
        VDF_TRY(pack_side(ctx, d_refs, perm, n_ref, false, ctx->ref_rows, reinterpret_cast<uint32_t*>(pl.stats.as<unsigned long long>() + 3)));
        ref_window_kernel<<<(r_pad + 255) / 256, 256, 0, ctx->stream>>>(cand.d_dur, (uint32_t)n_cand, d_ref_dur, perm, (uint32_t)n_ref, r_pad,
                                                                        pl.row_lo.as<uint32_t>(), pl.row_hi.as<uint32_t>());
        VDF_LAUNCHED(ctx);
        // in multi-GPU use every rank holds a different candidate slice and evaluates ALL of its tiles
        const uint32_t world = ctx->world, rank = ctx->rank;
        ctx->world = 1, ctx->rank = 0;
        const int rc = plan_finish(ctx, pl, TR, TC, &cand, 3);
        ctx->world = world, ctx->rank = rank;
        VDF_TRY(rc);
    }
    return run_plan(ctx, pl, ctx->ref_rows, cols, perm, cand_base, tol, d_keys_out, capacity, n_out);
}

int search_refs_device(vdf_ctx* ctx, const uint64_t* d_cand, const uint32_t* d_cand_dur, uint64_t n_cand,
                       uint64_t cand_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                       uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    *n_out = 0;
    if ((n_cand == 0 || n_ref == 0) && !ctx->exchange) return VDF_OK;
    if (n_cand >= 0xFFFFFF00ull || n_ref >= 0xFFFFFF00ull || cand_base + n_cand > 0xFFFFFFFFull) {
        ctx->err = "indices must fit in 32 bits";
        return VDF_ERR_INVALID;
    }
    VDF_TRY(table_prepare(ctx, ctx->tmp_cand, d_cand, nullptr, d_cand_dur, n_cand, false, true));
    return table_search_refs(ctx, ctx->tmp_cand, cand_base, d_refs, d_ref_dur, n_ref, tol, d_keys_out, capacity, n_out);
}

__global__ void window_pairs_kernel(const uint32_t* __restrict__ dur, uint32_t n, unsigned long long* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < n) {
        uint32_t thresh = f64_as_u32((double)dur[i] * 1.1);
        v = upper_bound_u32(dur, n, thresh) - (i + 1);
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

int self_window_pairs(vdf_ctx* ctx, const uint32_t* d_dur, uint64_t n, uint64_t* pairs_out) {
    *pairs_out = 0;
    if (n == 0) return VDF_OK;
    VDF_ALLOC(ctx, ctx->misc.ensure(64));
    VDF_CUDA(ctx, cudaMemsetAsync(ctx->misc.p, 0, 8, ctx->stream));
    window_pairs_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(d_dur, (uint32_t)n,
                                                                             ctx->misc.as<unsigned long long>());
    VDF_LAUNCHED(ctx);
    unsigned long long v = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(&v, ctx->misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *pairs_out = v;
    return VDF_OK;
}

}  // namespace vdf
