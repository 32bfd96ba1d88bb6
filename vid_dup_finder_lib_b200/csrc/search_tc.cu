// search_tc.cu -- tensor-core Hamming search: search_variant 5 (kind::i8) and 6 (kind::mxf4).
//
// hamming(a, b) = pc(a) + pc(b) - 2 * <a, b>: the all-pairs comparison is an exact contraction over the K = 1024 stored bits
// of two hashes that runs on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).  A pair matches
// iff   2 * dot(i, j) - pc(j) >= pc(i) - tol.   Results are bit-identical to the XOR+POPC kernels of search.cu (integers
// throughout) and are validated against the same oracle tests.
//
// HBM holds only PACKED tiles (bits); they are expanded to tensor-core operands inside the kernels.  Work unit = a CTA pair
// (cluster of 2, tcgen05.mma.cta_group::2, M = 256) x a chunk of column super-tiles.  Round 1 also carried variants 3 and 4
// (byte-expanded tiles in HBM, one CTA / CTA pairs): they needed 1 GB per million hashes, were bound by L2 -> SM bandwidth and
// were strictly dominated by 5 and 6; they were removed in round 2 (the measurements stay in profiles/r01_*).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vdf {

constexpr int kTcChunkBytes = kTile * 128;        // one K-chunk of one tile: 16 KB
constexpr int kTcTileBytes = 8 * kTcChunkBytes;   // expanded tile: 128 KB

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(tc_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tc_smem_u32(bar))
                 : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), LBO in
// [16,30) (ignored for swizzled K-major, 1), SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// one lane of a converged warp (the same lane every time): the canonical predicate for issuing tcgen05.mma / commit.
// Keeping the MMA warp converged and electing here lets ptxas keep barrier addresses and descriptors in uniform
// registers; a `lane == 0` branch around the loop costs a vote loop (ELECT / BRA.U.ANY) per UTCIMMA instead.
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define TC_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
// 32 TMEM lanes (one per thread of the warp) x 64 consecutive 32-bit columns
__device__ __forceinline__ void tc_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,"
        "%62,%63}, [%64];"
        : TC_R8(v, 0), TC_R8(v, 8), TC_R8(v, 16), TC_R8(v, 24), TC_R8(v, 32), TC_R8(v, 40), TC_R8(v, 48), TC_R8(v, 56)
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
    const uint8_t* row_exp;
    const uint8_t* col_exp;
    const uint32_t* row_pc;
    const uint32_t* col_pc;
    const uint32_t* col_pcmin;  // variant 5: min of col_pc over each group of 64 columns
    const uint32_t* row_lo;
    const uint32_t* row_hi;
    const uint32_t* row_id;
    const uint2* tile_range;  // in 128-hash column tiles
    uint64_t* keys;
    unsigned long long* counter;
    uint64_t capacity;
    uint64_t col_base;
    uint32_t chunk;  // column SUPER-tiles (256 hashes) per CTA
    uint32_t tol;
    uint32_t rank, world;
    const uint32_t* unit_off;  // variant 5: exclusive scan of the work units each row pair owns on this rank
    uint32_t n_pairs;
    const uint64_t* unit_list;  // variant 6: (chunk << 32 | row pair) of every unit this rank owns, chunk-major
    PeerPtrs peers;             // variant 6: peers.world > 0 => matches go to every rank's exchange buffer (common.cuh)
};

// ------------------------------------------------------------------------------------------------ CTA pairs
// A CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256) owns a row super-tile of 256 hashes (each CTA keeps its own
// 128 rows) and each CTA stages only ITS half of every column super-tile, so the L2 -> SM operand traffic and the shared-
// memory operand reads per pair of hashes are half of a single CTA's.
//   * the leader's tcgen05.commit multicasts to empty[stage] / acc_full[buf] of BOTH CTAs;
//   * acc_empty[buf] lives in the leader and counts the 8 epilogue warps of the pair.
// Column chunks are absolute (chunk c = super-tiles [c*chunk, (c+1)*chunk)), so that the pairs resident at the same
// time stream the same column tiles through L2.
constexpr int kTc2Stages = 4;
constexpr int kTc2StageBytes = kTcChunkBytes;  // this CTA's 128 columns x 128 B
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = unsigned 8 bit (0), both K-major,
// N = 256 (>> 3 at bit 17), M = 256 (>> 4 at bit 24)
constexpr uint32_t kTc2Idesc = (2u << 4) | (32u << 17) | (16u << 24);

__device__ __forceinline__ uint32_t tc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  CTA-scope semantics on purpose:
// everything these barriers order is async-proxy traffic (bulk copies -> shared memory -> UMMA, tensor memory), and a
// cluster-scope acquire makes ptxas invalidate the whole L1 (CCTL.IVALL) after every wait -- 41 % of all stall samples in
// the first capture of the first CTA-pair kernel (profiles/r01_hamming_v4_ncu.txt)
__device__ __forceinline__ void tc_mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(tc_smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kTc2Idesc), "r"(accumulate)
        : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in both CTAs
__device__ __forceinline__ void tc2_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            tc_smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// ------------------------------------------------------------------------------------------------ CTA pair, packed operands
// search_variant 5: the 1024 stored bits of a hash as 1024 bytes, u8 x u8 -> s32 (kind::i8), with the bit -> byte expansion
// INSIDE the kernel.  HBM keeps only packed tiles (pk[tile][K-chunk 0..7][hash 0..127][4 x u32], 16 KB per 128 hashes), so
// the operand traffic from L2 is 1/8 of a byte-expanded table's and no such copy (1 GB per million hashes) exists.  Four expander warps turn each K-chunk of
// this CTA's column tile (128 hashes x 128 bits, read from a packed tile that a bulk copy landed in shared memory a whole
// tile ahead) into the 128 x 128 B swizzled UMMA operand stage, make the writes visible to the async proxy
// (fence.proxy.async) and arrive on the LEADER's full[stage] (8 arrivals: 4 warps x 2 CTAs).  A stage is refilled from
// shared memory instead of from L2, so the ring turn-around drops from ~2000 to a few hundred cycles.
//
// Operand values: bit m (0..7) of every byte group becomes the byte  bit << m  in B and  bit << (7 - m)  in A, so every
// common bit contributes exactly 2^7 to the u8 x u8 -> s32 dot product (<= 2^17 in total) and the B expansion is one
// PRMT (replicate a source byte) + two LOP3 (mask) per eight output bytes.  dot = acc >> 7.
constexpr int kTc5Threads = 320;  // warp 0 bulk-copy producer, 1 MMA, 2-5 epilogue, 6-9 expanders
constexpr int kTc5PackedBufs = 2;
constexpr size_t kTc5Smem =
    (size_t)kTcTileBytes + kTc2Stages * kTc2StageBytes + kTc5PackedBufs * (size_t)kTileWords * 4 + 1024 /*align*/ + 256;

__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// lane -> (row-in-8, word): lanes of a quarter-warp cover rows 2q, 2q+1 x words 0..3, so that their 16-byte stores hit
// eight different 16-byte bank groups of the 128 B-swizzled line pair; the packed reads of a warp are 128 contiguous bytes
__device__ __forceinline__ int tc5_lane_row(int lane) { return ((lane >> 3) << 1) | ((lane >> 2) & 1); }

// B operand: one packed u32 -> 32 bytes (bit m of byte group -> bit << m)
__device__ __forceinline__ void tc5_expand_b(uint32_t bits, uint8_t* line, int w, int rsw) {
    uint4 lo, hi;
    const uint32_t b0 = __byte_perm(bits, 0, 0x0000), b1 = __byte_perm(bits, 0, 0x1111);
    const uint32_t b2 = __byte_perm(bits, 0, 0x2222), b3 = __byte_perm(bits, 0, 0x3333);
    lo.x = b0 & 0x08040201u, lo.y = b0 & 0x80402010u, lo.z = b1 & 0x08040201u, lo.w = b1 & 0x80402010u;
    hi.x = b2 & 0x08040201u, hi.y = b2 & 0x80402010u, hi.z = b3 & 0x08040201u, hi.w = b3 & 0x80402010u;
    *reinterpret_cast<uint4*>(line + (((2 * w) ^ rsw) << 4)) = lo;
    *reinterpret_cast<uint4*>(line + (((2 * w + 1) ^ rsw) << 4)) = hi;
}
// A operand: bit m of byte group -> bit << (7 - m)   (once per work unit: speed is irrelevant)
__device__ __forceinline__ uint32_t tc5_a_word(uint32_t nib, bool upper) {
    const uint32_t s = (nib * 0x00204081u) & 0x01010101u;  // byte i = bit i of the nibble
    const uint32_t v = ((s & 0x00000001u) << 7) | ((s & 0x00000100u) << 6) | ((s & 0x00010000u) << 5) | ((s & 0x01000000u) << 4);
    return upper ? (v >> 4) : v;
}
__device__ __forceinline__ void tc5_expand_a(uint32_t bits, uint8_t* line, int w, int rsw) {
    uint4 lo, hi;
    lo.x = tc5_a_word(bits & 0xFu, false), lo.y = tc5_a_word((bits >> 4) & 0xFu, true);
    lo.z = tc5_a_word((bits >> 8) & 0xFu, false), lo.w = tc5_a_word((bits >> 12) & 0xFu, true);
    hi.x = tc5_a_word((bits >> 16) & 0xFu, false), hi.y = tc5_a_word((bits >> 20) & 0xFu, true);
    hi.z = tc5_a_word((bits >> 24) & 0xFu, false), hi.w = tc5_a_word(bits >> 28, true);
    *reinterpret_cast<uint4*>(line + (((2 * w) ^ rsw) << 4)) = lo;
    *reinterpret_cast<uint4*>(line + (((2 * w + 1) ^ rsw) << 4)) = hi;
}

// Work units of variant 5.  A unit = (row pair P, absolute chunk c of column super-tiles); pair P needs the chunks that
// intersect the union of its two row tiles' column ranges, and rank r owns those with (P + c) % world == r.  The grid
// holds exactly the owned units (an exclusive scan over the pairs maps blockIdx -> (P, k-th owned chunk)): a 2-D grid
// with early exits launches ~8x more clusters than it uses on 8 GPUs, and an empty cluster still costs about a
// microsecond of an SM pair.
__device__ __forceinline__ bool tc5_pair_chunks(const TcParams& p, uint32_t n_row_tiles, uint32_t P, uint32_t* st_lo,
                                                uint32_t* st_hi, uint32_t* c_first, uint32_t* n_owned) {
    const uint2 r0 = p.tile_range[2 * P];
    const uint2 r1 = (2 * P + 1 < n_row_tiles) ? p.tile_range[2 * P + 1] : make_uint2(0, 0);
    uint32_t t_lo = 0xFFFFFFFFu, t_hi = 0;
    if (r0.x < r0.y) t_lo = r0.x, t_hi = r0.y;
    if (r1.x < r1.y) t_lo = min(t_lo, r1.x), t_hi = max(t_hi, r1.y);
    *n_owned = 0;
    if (t_lo >= t_hi) return false;
    *st_lo = t_lo / 2, *st_hi = (t_hi + 1) / 2;  // column super-tiles [st_lo, st_hi)
    const uint32_t c_lo = *st_lo / p.chunk, c_hi = (*st_hi - 1) / p.chunk;
    const uint32_t skip = (p.rank + p.world - (P + c_lo) % p.world) % p.world;  // first owned chunk at or after c_lo
    *c_first = c_lo + skip;
    if (*c_first > c_hi) return false;
    *n_owned = (c_hi - *c_first) / p.world + 1;
    return true;
}

__global__ void tc5_units_kernel(const TcParams p, uint32_t n_row_tiles, uint32_t* __restrict__ cnt) {
    const uint32_t P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P > p.n_pairs) return;
    uint32_t a, b, c, n = 0;
    if (P < p.n_pairs) tc5_pair_chunks(p, n_row_tiles, P, &a, &b, &c, &n);
    cnt[P] = n;  // cnt[n_pairs] = 0: the scan's last element is the total
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc5Threads, 1)
    hamming_tc5_kernel(const TcParams p, uint32_t n_row_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles (offset arithmetic keeps the pointers in the shared space)
    uint8_t* base = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;
    uint8_t* sB = base + kTcTileBytes;
    uint32_t* sP = reinterpret_cast<uint32_t*>(sB + kTc2Stages * kTc2StageBytes);  // [2][8 kc][128 rows][4] packed tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kTc5PackedBufs * kTileWords);
    uint64_t* full = bars;                     // [kTc2Stages]  leader: 8 expander-warp arrivals
    uint64_t* empty = bars + kTc2Stages;       // [kTc2Stages]  commit multicast
    uint64_t* pfull = bars + 2 * kTc2Stages;   // [2] packed tile landed
    uint64_t* pempty = pfull + 2;              // [2] 4 expander warps done with it
    uint64_t* acc_full = pempty + 2;           // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]  (leader's copy is the live one)
    uint64_t* a_full = acc_empty + 2;          // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

    const uint32_t cr = tc_cluster_rank();
    // blockIdx.x >> 1 = the unit; the pair that owns it = the last P with unit_off[P] <= unit
    const uint32_t unit = blockIdx.x >> 1;
    uint32_t lo = 0, hi = p.n_pairs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(p.unit_off + mid) <= unit) lo = mid;
        else hi = mid;
    }
    const uint32_t P = lo;
    uint32_t st_lo, st_hi, c_first, n_owned;
    if (!tc5_pair_chunks(p, n_row_tiles, P, &st_lo, &st_hi, &c_first, &n_owned)) return;  // cannot happen for a listed unit
    const uint32_t c = c_first + (unit - __ldg(p.unit_off + P)) * p.world;
    const uint32_t st0 = max(st_lo, c * p.chunk), st1 = min(st_hi, (c + 1) * p.chunk);
    if (st0 >= st1) return;
    const uint32_t n_st = st1 - st0;
    const uint32_t I = 2 * P + cr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t* row_tiles = reinterpret_cast<const uint32_t*>(p.row_exp);
    const uint32_t* col_tiles = reinterpret_cast<const uint32_t*>(p.col_exp);

    if (tid == 0) {
        for (int s = 0; s < kTc2Stages; ++s) tc_mbar_init(&full[s], 8), tc_mbar_init(&empty[s], 1);
        for (int b = 0; b < 2; ++b) {
            tc_mbar_init(&pfull[b], 1), tc_mbar_init(&pempty[b], 4);
            tc_mbar_init(&acc_full[b], 1), tc_mbar_init(&acc_empty[b], 8);
        }
        tc_mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // this CTA's 128 packed rows land in packed buffer 0 and are expanded by everybody below
        tc_mbar_expect_tx(a_full, kTileWords * 4);
        tc_bulk_g2s(sP, row_tiles + (size_t)I * kTileWords, kTileWords * 4, a_full);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();  // barrier init visible to this CTA's waiters
    tc_mbar_wait(a_full, 0);
    {
        const int w = lane & 3, r8 = tc5_lane_row(lane);
        for (int item = warp; item < 8 * 16; item += kTc5Threads / 32) {  // (K-chunk, 8-row group)
            const int kc = item >> 4, row = (item & 15) * 8 + r8;
            tc5_expand_a(sP[(kc * kTile + row) * 4 + w], sA + kc * kTcChunkBytes + row * 128, w, r8);
        }
    }
    tc_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();  // both A halves expanded, both CTAs' barriers initialised, tensor memory allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== bulk-copy producer: this CTA's packed column tile of every super-tile, one tile ahead
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t pb = s & 1;
                tc_mbar_wait(&pempty[pb], ((s >> 1) & 1) ^ 1);
                tc_mbar_expect_tx(&pfull[pb], kTileWords * 4);
                tc_bulk_g2s(sP + pb * kTileWords, col_tiles + (size_t)(2 * (st0 + s) + cr) * kTileWords, kTileWords * 4, &pfull[pb]);
            }
        }
    } else if (warp == 1) {
        if (cr == 0) {  // ===== MMA issuer (leader CTA only): converged warp, one elected lane issues
            const uint64_t a_desc = tc_desc(tc_smem_u32(sA)), b_desc = tc_desc(tc_smem_u32(sB));
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t buf = s & 1;
                tc_mbar_wait(&acc_empty[buf], ((s >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {  // it = 8 s + kc: stage and parity are compile-time (8 = 2 x kTc2Stages)
                    tc_mbar_wait(&full[kc & 3], (kc >> 2) & 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc2_mma(d_tmem, a_desc + ((kc * kTcChunkBytes + ks * 32) >> 4),
                                    b_desc + (((kc & 3) * kTc2StageBytes + ks * 32) >> 4), (kc | ks) != 0);
                        tc2_commit(&empty[kc & 3]);
                        if (kc == 7) tc2_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 6) {  // ===== expanders: warp e owns rows 32e .. 32e+31 of every stage
        const int w = lane & 3, r8 = tc5_lane_row(lane);
        const int row = (warp - 6) * 32 + r8;  // + 8 g
        uint32_t it = 0;
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t pb = s & 1;
            tc_mbar_wait(&pfull[pb], (s >> 1) & 1);
            const uint32_t* packed = sP + pb * kTileWords + row * 4 + w;
            for (int kc = 0; kc < 8; ++kc, ++it) {
                uint32_t bits[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) bits[g] = packed[(kc * kTile + g * 8) * 4];
                const uint32_t stage = it % kTc2Stages;
                tc_mbar_wait(&empty[stage], ((it / kTc2Stages) & 1) ^ 1);
                uint8_t* line = sB + stage * kTc2StageBytes + row * 128;
#pragma unroll
                for (int g = 0; g < 4; ++g) tc5_expand_b(bits[g], line + g * 8 * 128, w, r8);
                tc_fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive_remote(&full[stage], 0);
            }
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&pempty[pb]);
        }
    } else {  // ===== epilogue (warps 2-5 of both CTAs): warp w reads TMEM lanes 32*(w%4) .. +31 of its own CTA
        const uint32_t quarter = warp & 3;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t gi = I * kTile + row;
        const bool live = I < n_row_tiles;
        const int thr = live ? (int)p.row_pc[gi] - (int)p.tol : 0x7FFFFFFF;
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t buf = s & 1;
            tc_mbar_wait(&acc_full[buf], (s >> 1) & 1);
            tc_fence_after();
            const uint32_t col_first = (st0 + s) * 256;
            const uint32_t* pcmin = p.col_pcmin + (col_first >> 6);
            for (int q = 0; q < 4; ++q) {
                uint32_t v[64];
                // a pair (i, j) matches iff 2 dot - pc(j) >= pc(i) - tol, with acc = 128 dot.  Screen 64 columns at once:
                // 2 max(dot) - min(pc(j)) bounds every column's left-hand side from above, so the common case costs one
                // 3-input max per two accumulators and no per-column load
                const int floor_pc = (int)__ldg(pcmin + q);
                __syncwarp();
                tc_ld64(tmem_base + buf * 256 + q * 64 + ((quarter * 32) << 16), v);
                uint32_t best = 0;
#pragma unroll
                for (int k = 0; k < 64; k += 2) best = max(best, max(v[k], v[k + 1]));
                if ((int)(best >> 6) - floor_pc >= thr) {  // rare: exact test of the 64 columns
                    const uint4* pcj = reinterpret_cast<const uint4*>(p.col_pc + col_first + q * 64);
                    uint64_t mask = 0;
#pragma unroll
                    for (int k4 = 0; k4 < 16; ++k4) {
                        const uint4 pj = __ldg(pcj + k4);
                        mask |= (uint64_t)((int)(v[4 * k4 + 0] >> 6) - (int)pj.x >= thr) << (4 * k4 + 0);
                        mask |= (uint64_t)((int)(v[4 * k4 + 1] >> 6) - (int)pj.y >= thr) << (4 * k4 + 1);
                        mask |= (uint64_t)((int)(v[4 * k4 + 2] >> 6) - (int)pj.z >= thr) << (4 * k4 + 2);
                        mask |= (uint64_t)((int)(v[4 * k4 + 3] >> 6) - (int)pj.w >= thr) << (4 * k4 + 3);
                    }
                    while (mask) {
                        const int k = __ffsll((long long)mask) - 1;
                        mask &= mask - 1;
                        const uint32_t gj = col_first + q * 64 + k;
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive_remote(&acc_empty[buf], 0);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// hashes [n][32] u32 -> pk[tile][K-chunk][hash][4 u32] + per-hash popcounts (the epilogue's pc(i), pc(j)); one thread per
// hash; tiles beyond n are zero.  perm (optional) gathers rows.
__global__ void __launch_bounds__(kTile) tc5_pack_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm,
                                                         uint64_t n, uint32_t* __restrict__ pk, uint32_t* __restrict__ pc,
                                                         uint32_t* __restrict__ pcmin64) {
    __shared__ uint32_t wmin[4];
    const uint64_t g = (uint64_t)blockIdx.x * kTile + threadIdx.x;
    uint4* out = reinterpret_cast<uint4*>(pk + (size_t)blockIdx.x * kTileWords) + threadIdx.x;
    uint32_t c = 0;
    const uint4* src = g < n ? reinterpret_cast<const uint4*>(in) + (perm ? perm[g] : g) * 8 : nullptr;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src) v = src[kc];
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        out[kc * kTile] = v;
    }
    pc[g] = c;
    const uint32_t m = __reduce_min_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 2) pcmin64[(size_t)blockIdx.x * 2 + threadIdx.x] = min(wmin[2 * threadIdx.x], wmin[2 * threadIdx.x + 1]);
}

int tc5_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, Packed& out) {
    const uint32_t T = (uint32_t)((n + kTile - 1) / kTile);
    const uint32_t T2 = (T + 1) & ~1u;  // CTA pairs read tile pairs: one zero tile of padding when T is odd
    VDF_ALLOC(ctx, out.tiles.ensure((size_t)T2 * kTileWords * 4));
    VDF_ALLOC(ctx, out.pc.ensure((size_t)T2 * kTile * 4));
    VDF_ALLOC(ctx, out.pcmin.ensure((size_t)T2 * 2 * 4));
    tc5_pack_kernel<<<T2, kTile, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_hash), perm, n, out.tiles.as<uint32_t>(),
                                                  out.pc.as<uint32_t>(), out.pcmin.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    out.n = n, out.variant = 5, out.as_columns = false;
    return VDF_OK;
}

// ------------------------------------------------------------------------------------------------ CTA pair, 4-bit operands
// search_variant 6 (default): variant 5 with the bits expanded to e2m1 nibbles and tcgen05.mma kind::mxf4 (block-scaled,
// K = 64 per instruction, fp32 accumulators: exact, every partial sum is a small integer), which issues at twice the
// kind::i8 rate.  Block scales are the constant 1.0 (UE8M0 0x7F): the tensor-memory columns the scale factors live in are
// filled with 0x7F bytes once, so their layout does not matter.  Those columns have to come out of the accumulators' 512,
// hence super-tiles of 192 columns (two 192-column accumulators at [0,192) and [192,384), scale factors behind them) and
// column tiles of 96 hashes.  Operands are half the size of variant 5's (A: 64 KB per CTA), which pays for an 8-stage ring.
// K order: nibble j of output word m of a unit holds bit 4j + m of the packed word -- any permutation of K is fine as long
// as A and B use the same one.
//
// The fold (kFold, round 2).  A real VideoHash never sets bits 1000..1023 (dct_3d.rs:55-66 writes 1000 bits), so 24 of the
// K = 1024 positions are free.  When the pack kernels saw those bits zero in BOTH operands, the kernel puts
//     rows:    every hash bit at 2.0 (e2m1 0x4), and the constants 6.0 x 23, 2.0 x 1 in the free positions;
//     columns: every hash bit at 1.0 (e2m1 0x2), and kFoldC - pc(j) spelled in e2m1 digits in the free positions
//              (3 x an integer from 23 signed digits of {0, .5, 1, 1.5, 2, 3, 4, 6} against the 6.0s, the remainder 0..2 as
//              {0, .5, 1} against the 2.0; one precomputed 16-byte operand unit per column, appended to every packed tile)
// so the contraction returns   acc = 2 dot(i, j) - pc(j) + kFoldC   EXACTLY (integers below 2^12) and
//     hamming(i, j) <= tol   <=>   acc >= kFoldC + pc(i) - tol.
// The maximum over 64 accumulators is then an exact screen at ANY tolerance and the epilogue loads no popcounts: the round-1
// screen 2 max(dot) - min(pc(j)) stopped rejecting random hashes past tol ~0.39 (236 ms per 1 M launch at 0.40 against 120).
// Tables with a pad bit set anywhere (the test helpers of video_hash.rs:265-280 make such hashes) take the round-1 epilogue.
constexpr int kT6Cols = 96;                          // hashes per column tile (one CTA's half of a super-tile)
constexpr int kT6Chunk = 128;                        // bytes per row per K-chunk = 256 bits
constexpr int kT6ABytes = 4 * kTile * kT6Chunk;      // 64 KB
constexpr int kT6StageBytes = kT6Cols * kT6Chunk;    // 12 KB
constexpr int kT6Stages = 8;
constexpr int kT6Expanders = 4;                      // expander warps (24 four-row groups per stage: 6 each)
constexpr int kTc6Threads = (6 + kT6Expanders) * 32; // warp 0 producer, 1 MMA, 2-5 epilogue, 6.. expanders
constexpr int kT6PackedBytes = kT6Cols * 128;        // packed bits of a column tile: 96 x 128 B = 12 KB; packed row tile: 16 KB
constexpr int kT6FoldBytes = kT6Cols * 16;           // + one fold unit per column
constexpr int kT6ColTileBytes = kT6PackedBytes + kT6FoldBytes;  // 13.5 KB per column tile in HBM, one bulk copy
constexpr int kFoldC = 800;                          // pc(j) <= 1000: kFoldC - pc(j) in [-200, 800], |.| <= 3 * 276 + 2
// all-shared form: A 64 KB + ring + 2 packed slots; kATmem form: A chunk 3 (16 KB) + ring + 4 packed slots (smaller)
constexpr size_t kTc6Smem = (size_t)kT6ABytes + kT6Stages * kT6StageBytes + 2 * kTileWords * 4 /* packed A, then 2 packed column tiles */ + 1024 + 256;
// block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled): a/b format E2M1 = 1 at [7,10) / [10,13),
// K-major, N >> 3 at [17,23), scale format UE8M0 = 1 at bit 23, M >> 4 at [24,29), scale-factor ids 0, K = 64
constexpr uint32_t kTc6Idesc = (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | (1u << 23) | ((256u >> 4) << 24);

__device__ __forceinline__ void tc6_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kTc6Idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
}
// the same with operand A read from tensor memory (8 columns = 64 e2m1 values per lane and instruction)
__device__ __forceinline__ void tc6_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(kTc6Idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
}
// one packed u32 -> one 16-byte unit of e2m1 nibbles: 0x2 = 1.0 per set bit (columns; rows without the fold) ...
__device__ __forceinline__ uint4 tc6_expand(uint32_t w) {
    return make_uint4((w << 1) & 0x22222222u, w & 0x22222222u, (w >> 1) & 0x22222222u, (w >> 2) & 0x22222222u);
}
// ... or 0x4 = 2.0 per set bit (rows with the fold)
template <bool kFold>
__device__ __forceinline__ uint4 tc6_expand_row(uint32_t w) {
    if (!kFold) return tc6_expand(w);
    return make_uint4((w << 2) & 0x44444444u, (w << 1) & 0x44444444u, w & 0x44444444u, (w >> 1) & 0x44444444u);
}
// Free K positions: the unit of hash word 31 (u32) holds hash bits 992..1023; nibble j of output word m is bit 4j + m, so the
// 24 pad positions are nibbles 2..7 of all four words.  Rows carry 6.0 (0x7) in 23 of them and 2.0 (0x4) in the last.
constexpr uint32_t kFoldRowPad = 0x77777700u, kFoldRowPadLast = 0x47777700u;

// kFoldC - pc spelled in e2m1 digits (see the head of this section): out = the column's unit for hash word 31
__device__ __forceinline__ uint4 tc6_fold_unit(uint32_t w31, uint32_t pc) {
    const int T = kFoldC - (int)pc;
    int r = T % 3;
    if (r < 0) r += 3;
    const int Q = (T - r) / 3;  // sum over 23 digits of 2 * digit
    const uint32_t sign = Q < 0 ? 8u : 0u;
    int q = Q < 0 ? -Q : Q;
    const uint4 e = tc6_expand(w31);
    uint32_t word[4] = {e.x & 0xFFu, e.y & 0xFFu, e.z & 0xFFu, e.w & 0xFFu};
#pragma unroll
    for (int d = 0; d < 23; ++d) {
        // largest doubled digit value <= q from {12, 8, 6, 4, 3, 2, 1}; its e2m1 code: 1..4 -> 1..4, 6 -> 5, 8 -> 6, 12 -> 7
        const int v = q >= 12 ? 12 : q >= 8 ? 8 : q >= 6 ? 6 : q >= 4 ? 4 : q;  // q < 4: q itself (0..3)
        const uint32_t code = v <= 4 ? (uint32_t)v : v == 6 ? 5u : v == 8 ? 6u : 7u;
        q -= v;
        word[d / 6] |= (v ? (code | sign) : 0u) << (4 * (2 + d % 6));
    }
    word[3] |= (uint32_t)r << 28;  // digit 23 = (m 3, j 7): 0, 0.5 or 1.0 against the row's 2.0
    return make_uint4(word[0], word[1], word[2], word[3]);
}
#define TC_RI8(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(
            taddr),
        TC_RI8(v, 0), TC_RI8(v, 8), TC_RI8(v, 16), TC_RI8(v, 24)
        : "memory");
}

// hashes [n][32] u32 -> pk[tile][K-chunk 0..3][row 0..T-1][8 x u32] (+ [row][4 x u32] fold units for column tiles) + pc;
// pads_or |= bits 1000..1023 of every hash.  One thread per hash; tiles beyond n are zero.  perm (optional) gathers rows.
template <int T, bool kCols>
__global__ void __launch_bounds__(128) tc6_pack_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, uint64_t n,
                                                       uint32_t* __restrict__ pk, uint32_t* __restrict__ pc, uint32_t* __restrict__ pads_or) {
    if ((int)threadIdx.x >= T) return;  // T is a multiple of 32: whole warps leave
    constexpr size_t kTileU32 = kCols ? kT6ColTileBytes / 4 : (size_t)T * 32;
    const uint64_t g = (uint64_t)blockIdx.x * T + threadIdx.x;
    uint32_t* tile = pk + (size_t)blockIdx.x * kTileU32;
    uint4* out = reinterpret_cast<uint4*>(tile) + threadIdx.x * 2;
    uint32_t c = 0, w31 = 0;
    const uint4* src = g < n ? reinterpret_cast<const uint4*>(in) + (perm ? perm[g] : g) * 8 : nullptr;
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // q = 16-byte piece of the hash; K-chunk kc = q / 2 holds pieces 2kc, 2kc+1
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src) v = src[q];
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        out[(q >> 1) * T * 2 + (q & 1)] = v;
        if (q == 7) w31 = v.w;
    }
    pc[g] = c;
    if (kCols) reinterpret_cast<uint4*>(tile + (size_t)T * 32)[threadIdx.x] = tc6_fold_unit(w31, c);
    const uint32_t pad = __reduce_or_sync(0xffffffffu, w31 >> 8);
    if ((threadIdx.x & 31) == 0 && pad) atomicOr(pads_or, pad);
}
__global__ void pcmin64_kernel(const uint32_t* __restrict__ pc, uint64_t n_groups, uint32_t* __restrict__ pcmin) {
    const uint64_t g = blockIdx.x * (uint64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t m = __reduce_min_sync(0xffffffffu, min(pc[g * 64 + lane], pc[g * 64 + 32 + lane]));
    if (lane == 0) pcmin[g] = m;
}

__device__ __forceinline__ bool tc6_pair_chunks(const TcParams& p, uint32_t n_row_tiles, uint32_t P, uint32_t* st_lo,
                                                uint32_t* st_hi, uint32_t* c_first, uint32_t* n_owned) {
    const uint2 r0 = p.tile_range[2 * P];
    const uint2 r1 = (2 * P + 1 < n_row_tiles) ? p.tile_range[2 * P + 1] : make_uint2(0, 0);
    uint32_t t_lo = 0xFFFFFFFFu, t_hi = 0;
    if (r0.x < r0.y) t_lo = r0.x, t_hi = r0.y;
    if (r1.x < r1.y) t_lo = min(t_lo, r1.x), t_hi = max(t_hi, r1.y);
    *n_owned = 0;
    if (t_lo >= t_hi) return false;
    // tile_range counts 128-hash tiles; super-tiles here are 192 columns
    *st_lo = (uint32_t)(((uint64_t)t_lo * kTile) / (2 * kT6Cols));
    *st_hi = (uint32_t)(((uint64_t)t_hi * kTile + 2 * kT6Cols - 1) / (2 * kT6Cols));
    const uint32_t c_lo = *st_lo / p.chunk, c_hi = (*st_hi - 1) / p.chunk;
    const uint32_t skip = (p.rank + p.world - (P + c_lo) % p.world) % p.world;
    *c_first = c_lo + skip;
    if (*c_first > c_hi) return false;
    *n_owned = (c_hi - *c_first) / p.world + 1;
    return true;
}
__global__ void tc6_units_kernel(const TcParams p, uint32_t n_row_tiles, uint32_t* __restrict__ cnt) {
    const uint32_t P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P > p.n_pairs) return;
    uint32_t a, b, c, n = 0;
    if (P < p.n_pairs) tc6_pair_chunks(p, n_row_tiles, P, &a, &b, &c, &n);
    cnt[P] = n;
}

// the units of row pair P, in P-major order; a stable sort on the chunk bits then makes the list chunk-major, so that the
// CTA pairs resident together walk the SAME column chunk (3 MB) and the packed column tiles are served by L2.  In P-major
// order 74 resident pairs stream 74 different chunks (230 MB > L2) and every launch re-reads the table from HBM ~600 times.
__global__ void tc6_unit_list_kernel(const TcParams p, uint32_t n_row_tiles, uint64_t* __restrict__ list) {
    const uint32_t P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= p.n_pairs) return;
    uint32_t a, b, c_first, n = 0;
    if (!tc6_pair_chunks(p, n_row_tiles, P, &a, &b, &c_first, &n)) return;
    uint64_t* out = list + p.unit_off[P];
    for (uint32_t k = 0; k < n; ++k) out[k] = ((uint64_t)(c_first + k * p.world) << 32) | P;
}

// Append the matches of one row (mask = its matching columns among the 64 starting at c0).  Called ONLY by the lanes that
// hold a match, so the common case (nothing within the tolerance) pays nothing: a warp-wide vote per 64 columns cost 14 % of
// the whole kernel (144 against 126 ms per 1 M launch: the epilogue has no slack).  The lanes that did get here form a
// coalesced group and share one atomic: at tolerances where a few per cent of all pairs match, per-match atomics on one
// address would set the pace instead.
__device__ __forceinline__ void tc6_emit(const TcParams& p, uint64_t mask, uint32_t c0, uint32_t gi, uint32_t win_lo, uint32_t win_hi) {
    {   // columns [win_lo, win_hi) of this row, as bits relative to c0
        const long long a = (long long)win_lo - (long long)c0, b = (long long)win_hi - (long long)c0;
        const uint64_t below_b = b >= 64 ? ~0ull : b <= 0 ? 0ull : ((1ull << b) - 1);
        const uint64_t below_a = a >= 64 ? ~0ull : a <= 0 ? 0ull : ((1ull << a) - 1);
        mask &= below_b & ~below_a;
    }
    const cg::coalesced_group g = cg::coalesced_threads();
    const uint32_t cnt = (uint32_t)__popcll(mask);
    const uint32_t before = cg::exclusive_scan(g, cnt, cg::plus<uint32_t>());
    const uint32_t total = g.shfl(before + cnt, g.size() - 1);
    if (total == 0) return;
    unsigned long long base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(p.counter, (unsigned long long)total);
    base = g.shfl(base, 0);
    unsigned long long slot = base + before;
    const uint64_t rid = p.row_id ? p.row_id[gi] : gi;
    while (mask) {
        const int k = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        const uint64_t key = (rid << 32) | (uint64_t)(c0 + k + p.col_base);
        if (p.peers.world) {  // fused exchange: this rank's segment of every rank's buffer, over NVLink
            if (slot < p.peers.seg_cap)
                for (uint32_t r = 0; r < p.peers.world; ++r) p.peers.keys(r, p.peers.rank)[slot] = key;
        } else if (slot < p.capacity) {
            p.keys[slot] = key;
        }
        ++slot;
    }
}

// tcgen05.ld of 64 accumulator columns without the wait: the load of the next group is in flight while this one is screened
__device__ __forceinline__ void tc_ld64_nowait(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,"
        "%62,%63}, [%64];"
        : TC_R8(v, 0), TC_R8(v, 8), TC_R8(v, 16), TC_R8(v, 24), TC_R8(v, 32), TC_R8(v, 40), TC_R8(v, 48), TC_R8(v, 56)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One row against 64 columns: -> the matching columns.
//   fold:    acc = 2 dot - pc(j) + kFoldC is exact, so the signed-int maximum of the 64 bit patterns against the bit pattern
//            of thr (> 0) is an exact screen (positive floats order like their bit patterns, negative ones are negative ints)
//   no fold: acc = dot; first the bound 2 max(dot) - min(pc(j)), then the exact maximum of 2 dot - pc(j) on the FMA pipe
//            (0x4B000000 | pc is the float 2^23 + pc: no int <-> float conversions, which run on the XU pipe at 16 per clock
//            and made the 1 M launch 2.5x slower at tolerance 0.4), then the per-column mask
template <bool kFold>
__device__ __forceinline__ uint64_t tc6_screen(const uint32_t (&v)[64], int thr, float thr_f, bool screen, int thr_bits, int floor_pc,
                                               const uint32_t* __restrict__ col_pc) {
    uint64_t mask = 0;
    if (kFold) {
        int best = (int)v[0];
#pragma unroll
        for (int k = 0; k < 64; k += 2) best = max(best, max((int)v[k], (int)v[k + 1]));
        if (!screen || best >= thr_bits) {  // only a row that really has a match among these 64 columns gets here
#pragma unroll
            for (int k = 0; k < 64; ++k) mask |= (uint64_t)(__uint_as_float(v[k]) >= thr_f) << k;
        }
    } else {
        uint32_t best = 0;  // accumulators are non-negative floats: their bit patterns order like the values
#pragma unroll
        for (int k = 0; k < 64; k += 2) best = max(best, max(v[k], v[k + 1]));
        if (2 * (int)__uint_as_float(best) - floor_pc >= thr) {
            const uint4* pcj = reinterpret_cast<const uint4*>(col_pc);
            float top = -3.0e38f;
#pragma unroll
            for (int k4 = 0; k4 < 16; ++k4) {
                const uint4 pj = __ldg(pcj + k4);
                top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 0]), 2.0f, -__uint_as_float(0x4B000000u | pj.x)));
                top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 1]), 2.0f, -__uint_as_float(0x4B000000u | pj.y)));
                top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 2]), 2.0f, -__uint_as_float(0x4B000000u | pj.z)));
                top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 3]), 2.0f, -__uint_as_float(0x4B000000u | pj.w)));
            }
            if (top >= thr_f) {  // a pair of this row with one of the 64 columns is within the tolerance
#pragma unroll
                for (int k4 = 0; k4 < 16; ++k4) {
                    const uint4 pj = __ldg(pcj + k4);
                    mask |= (uint64_t)(fmaf(__uint_as_float(v[4 * k4 + 0]), 2.0f, -__uint_as_float(0x4B000000u | pj.x)) >= thr_f) << (4 * k4 + 0);
                    mask |= (uint64_t)(fmaf(__uint_as_float(v[4 * k4 + 1]), 2.0f, -__uint_as_float(0x4B000000u | pj.y)) >= thr_f) << (4 * k4 + 1);
                    mask |= (uint64_t)(fmaf(__uint_as_float(v[4 * k4 + 2]), 2.0f, -__uint_as_float(0x4B000000u | pj.z)) >= thr_f) << (4 * k4 + 2);
                    mask |= (uint64_t)(fmaf(__uint_as_float(v[4 * k4 + 3]), 2.0f, -__uint_as_float(0x4B000000u | pj.w)) >= thr_f) << (4 * k4 + 3);
                }
            }
        }
    }
    return mask;
}

// kATmem: K-chunks 0-2 of the row operand live in tensor memory (columns [384, 480), written once per unit with tcgen05.st)
// and only chunk 3 in shared memory; scale factors in [480, 512).  The shared-memory data pipe is what bounds the all-smem
// form (ncu: 50 % tensor-core operand reads + 32 % expander stores at 86 % tensor-pipe activity); reading three quarters of
// A from tensor memory takes the operand reads from 7 KB to 4 KB per instruction.
template <bool kATmem, bool kFold>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc6Threads, 1)
    hamming_tc6_kernel(const TcParams p, uint32_t n_row_tiles, uint32_t n_col_st) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;                                                             // 4 K-chunks x 128 rows x 128 B
    constexpr int kPB = kATmem ? 4 : 2;  // packed column-tile buffers (16 KB slots; slot 0 holds the packed row tile first)
    constexpr int kASmem = kATmem ? kTile * kT6Chunk : kT6ABytes;  // only K-chunk 3 of A lives in shared memory with kATmem
    uint8_t* sB = base + kASmem;                                                    // ring: 8 x (96 rows x 128 B)
    uint32_t* sP = reinterpret_cast<uint32_t*>(sB + kT6Stages * kT6StageBytes);     // packed A tile (16 KB), then 2 x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sP) + kPB * kTileWords * 4);  // slots for column tiles
    uint64_t* full = bars;                      // [8]  leader: 8 expander-warp arrivals
    uint64_t* empty = bars + kT6Stages;         // [8]  commit multicast
    uint64_t* pfull = bars + 2 * kT6Stages;     // [kPB]
    uint64_t* pempty = pfull + kPB;             // [kPB]
    uint64_t* acc_full = pempty + kPB;          // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint64_t* a_full = acc_empty + 2;           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

    const uint32_t cr = tc_cluster_rank();
    const uint64_t unit = __ldg(p.unit_list + (blockIdx.x >> 1));
    const uint32_t P = (uint32_t)unit, c = (uint32_t)(unit >> 32);
    uint32_t st_lo, st_hi, c_first, n_owned;
    if (!tc6_pair_chunks(p, n_row_tiles, P, &st_lo, &st_hi, &c_first, &n_owned)) return;
    const uint32_t st0 = max(st_lo, c * p.chunk), st1 = min(min(st_hi, n_col_st), (c + 1) * p.chunk);
    if (st0 >= st1) return;
    const uint32_t n_st = st1 - st0;
    const uint32_t I = 2 * P + cr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t* row_tiles = reinterpret_cast<const uint32_t*>(p.row_exp);
    const uint32_t* col_tiles = reinterpret_cast<const uint32_t*>(p.col_exp);

    if (tid == 0) {
        for (int s = 0; s < kT6Stages; ++s) tc_mbar_init(&full[s], 2 * kT6Expanders), tc_mbar_init(&empty[s], 1);
        for (int b = 0; b < kPB; ++b) tc_mbar_init(&pfull[b], 1), tc_mbar_init(&pempty[b], kT6Expanders);
        for (int b = 0; b < 2; ++b) tc_mbar_init(&acc_full[b], 1), tc_mbar_init(&acc_empty[b], 8);
        tc_mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc_mbar_expect_tx(a_full, kTileWords * 4);
        tc_bulk_g2s(sP, row_tiles + (size_t)I * kTileWords, kTileWords * 4, a_full);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kSfCol = kATmem ? 480u : 384u;
    if (warp >= 2 && warp < 6) {  // constant block scales: 0x7F (UE8M0 1.0) in every byte of columns [kSfCol, 512), all lanes
        uint32_t ones[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) ones[k] = 0x7F7F7F7Fu;
        const uint32_t lanes = ((uint32_t)(warp & 3) * 32) << 16;
        for (uint32_t c = kSfCol; c < 512; c += 32) tc_st32(tmem_base + lanes + c, ones);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_mbar_wait(a_full, 0);
    if (kATmem && warp >= 2 && warp < 6) {  // A chunks 0-2 -> tensor memory: lane = row, column 384 + 32 kc + 4 w + m
        const uint32_t row = (uint32_t)(warp & 3) * 32 + lane;
        for (int kc = 0; kc < 3; ++kc) {
            const uint4* src = reinterpret_cast<const uint4*>(sP + (kc * kTile + row) * 8);
            const uint4 p0 = src[0], p1 = src[1];
            const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            uint32_t v[32];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint4 e = tc6_expand_row<kFold>(pw[w]);
                v[4 * w] = e.x, v[4 * w + 1] = e.y, v[4 * w + 2] = e.z, v[4 * w + 3] = e.w;
            }
            tc_st32(tmem_base + ((row & ~31u) << 16) + 384 + kc * 32, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    {   // A -> shared memory: K-chunks x 128 rows; a warp covers 4 rows x 8 packed words (one 128 B line per quarter-warp)
        const int w = lane & 7, r4 = lane >> 3;
        for (int item = warp + (kATmem ? 3 * 32 : 0); item < 4 * 32; item += kTc6Threads / 32) {  // (K-chunk, 4-row group)
            const int kc = item >> 5, row = (item & 31) * 4 + r4;
            uint4 e = tc6_expand_row<kFold>(sP[(kc * kTile + row) * 8 + w]);
            if (kFold && kc == 3 && w == 7) e.x |= kFoldRowPad, e.y |= kFoldRowPad, e.z |= kFoldRowPad, e.w |= kFoldRowPadLast;
            *reinterpret_cast<uint4*>(sA + (kATmem ? 0 : kc) * (kTile * kT6Chunk) + row * 128 + ((w ^ (row & 7)) << 4)) = e;
        }
    }
    tc_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();
    tc_fence_after();
    const uint32_t sfa = tmem_base + kSfCol, sfb = tmem_base + kSfCol + (kATmem ? 16u : 64u);

    if (warp == 0) {
        if (lane == 0) {  // ===== bulk-copy producer: this CTA's packed column tile (96 hashes + their fold units) of every super-tile
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t pb = s % kPB;
                tc_mbar_wait(&pempty[pb], ((s / kPB) & 1) ^ 1);
                tc_mbar_expect_tx(&pfull[pb], kT6ColTileBytes);
                tc_bulk_g2s(sP + pb * kTileWords, col_tiles + (size_t)(2 * (st0 + s) + cr) * (kT6ColTileBytes / 4), kT6ColTileBytes,
                            &pfull[pb]);
            }
        }
    } else if (warp == 1) {
        if (cr == 0) {  // ===== MMA issuer
            const uint64_t a_desc = tc_desc(tc_smem_u32(sA)), b_desc = tc_desc(tc_smem_u32(sB));
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t buf = s & 1;
                tc_mbar_wait(&acc_empty[buf], ((s >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 192;
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {  // it = 4 s + kc: stage = 4 (s & 1) + kc, parity = (s >> 1) & 1
                    const uint32_t stage = buf * 4 + kc;
                    tc_mbar_wait(&full[stage], (s >> 1) & 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (kATmem && kc < 3)
                                tc6_mma_ts(d_tmem, tmem_base + 384 + kc * 32 + ks * 8, b_desc + ((stage * kT6StageBytes + ks * 32) >> 4),
                                           sfa, sfb, (kc | ks) != 0);
                            else
                                tc6_mma(d_tmem, a_desc + (((kATmem ? 0 : kc) * (kTile * kT6Chunk) + ks * 32) >> 4),
                                        b_desc + ((stage * kT6StageBytes + ks * 32) >> 4), sfa, sfb, (kc | ks) != 0);
                        }
                        tc2_commit(&empty[stage]);
                        if (kc == 3) tc2_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 6) {  // ===== expanders: a warp writes 4 rows x 8 units per group, 6 groups per stage
        // Measured alternatives that did not help (1 M hashes, ms per launch): 6 or 8 warps on every stage (133.6 / 138.1
        // against 130.7), two teams of four warps on alternate stages (132.8 against 131.6), one fence.proxy.async per 2 / 4
        // stages (123.3 / 123.6 against 122.6), probing the next stage's `empty` barrier with mbarrier.test_wait before this
        // stage's stores (123.7 against 121.2).
        constexpr int kG = 24 / kT6Expanders;  // four-row groups per warp per stage
        const int w = lane & 7, r4 = lane >> 3;
        const int row0 = (warp - 6) * 4 + r4;  // + 4 * kT6Expanders * g
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t pb = s % kPB;
            tc_mbar_wait(&pfull[pb], (s / kPB) & 1);
            const uint32_t* packed = sP + pb * kTileWords + row0 * 8 + w;
            const uint4* fold_units = reinterpret_cast<const uint4*>(sP + pb * kTileWords + kT6PackedBytes / 4);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {  // unrolled: only K-chunk 3 carries the fold units, the other three stages pay nothing
                uint32_t bits[kG];
#pragma unroll
                for (int g = 0; g < kG; ++g) bits[g] = packed[(kc * kT6Cols + g * (4 * kT6Expanders)) * 8];
                const uint32_t it = 4 * s + kc;
                const uint32_t stage = it % kT6Stages;
                tc_mbar_wait(&empty[stage], ((it / kT6Stages) & 1) ^ 1);
                uint8_t* dst = sB + stage * kT6StageBytes;
#pragma unroll
                for (int g = 0; g < kG; ++g) {
                    const int row = row0 + 4 * kT6Expanders * g;
                    uint4 e = tc6_expand(bits[g]);
                    if (kFold && kc == 3 && w == 7) e = fold_units[row];  // the unit of hash word 31 comes precomputed with the column's digits
                    *reinterpret_cast<uint4*>(dst + row * 128 + ((w ^ (row & 7)) << 4)) = e;
                }
                tc_fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive_remote(&full[stage], 0);
            }
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&pempty[pb]);
        }
    } else {  // ===== epilogue
        const uint32_t quarter = warp & 3;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t gi = I * kTile + row;
        const bool live = I < n_row_tiles;
        const uint32_t win_lo = live ? __ldg(p.row_lo + gi) : 0u, win_hi = live ? __ldg(p.row_hi + gi) : 0u;
        // fold:    acc = 2 dot - pc(j) + kFoldC, a match iff acc >= kFoldC + pc(i) - tol
        // no fold: acc = dot,                    a match iff 2 dot - pc(j) >= pc(i) - tol
        const int thr = live ? (kFold ? kFoldC : 0) + (int)p.row_pc[gi] - (int)p.tol : 0x7FFFFFFF;
        const float thr_f = kFold ? (float)thr : (live ? (float)thr - 8388608.0f : 3.0e38f);  // no fold: 2 dot - pc(j) - 2^23 >= thr - 2^23, exact in fp32
        // positive floats order like their bit patterns taken as signed ints, and every negative one is a negative int:
        // with thr > 0 the signed-int maximum of 64 accumulators passes the bit pattern of thr iff some accumulator does
        const bool screen = thr > 0;
        const int thr_bits = __float_as_int((float)thr);
        const uint32_t lanes = (quarter * 32) << 16;
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t buf = s & 1;
            tc_mbar_wait(&acc_full[buf], (s >> 1) & 1);
            tc_fence_after();
            const uint32_t col_first = (st0 + s) * (2 * kT6Cols);
            const uint32_t* pcmin = p.col_pcmin + (col_first >> 6);
            const uint32_t acc = tmem_base + buf * 192 + lanes;
            // Three groups of 64 columns, software-pipelined: the tensor-memory load of group q+1 is in flight while group q is
            // screened (tcgen05.wait::ld waits for every load issued so far, so at most one is outstanding at a wait), and the
            // accumulator goes back to the MMA warp as soon as its last column is in registers, before the last screen.
            uint32_t va[64], vb[64];
            int floor0 = 0, floor1 = 0, floor2 = 0;
            if (!kFold) floor0 = (int)__ldg(pcmin), floor1 = (int)__ldg(pcmin + 1), floor2 = (int)__ldg(pcmin + 2);
            __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the rare emit path of the previous super-tile
            tc_ld64_nowait(acc, va);
            tc_wait_ld();
            tc_ld64_nowait(acc + 64, vb);
            uint64_t mask = tc6_screen<kFold>(va, thr, thr_f, screen, thr_bits, floor0, p.col_pc + col_first);
            if (mask) tc6_emit(p, mask, col_first, gi, win_lo, win_hi);
            __syncwarp();
            tc_wait_ld();
            tc_ld64_nowait(acc + 128, va);
            mask = tc6_screen<kFold>(vb, thr, thr_f, screen, thr_bits, floor1, p.col_pc + col_first + 64);
            if (mask) tc6_emit(p, mask, col_first + 64, gi, win_lo, win_hi);
            __syncwarp();
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) tc_mbar_arrive_remote(&acc_empty[buf], 0);
            mask = tc6_screen<kFold>(va, thr, thr_f, screen, thr_bits, floor2, p.col_pc + col_first + 128);
            if (mask) tc6_emit(p, mask, col_first + 128, gi, win_lo, win_hi);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
// rows: tiles of 128 hashes (+ pc); columns: tiles of 96 hashes + fold units, padded to whole super-tiles (+ pc, pcmin)
int tc6_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, bool as_columns, Packed& out, uint32_t* d_pads_or) {
    const uint32_t* in = reinterpret_cast<const uint32_t*>(d_hash);
    out.n = n, out.variant = 6, out.as_columns = as_columns;
    if (!as_columns) {
        const uint32_t T = (uint32_t)((n + kTile - 1) / kTile), T2 = (T + 1) & ~1u;
        VDF_ALLOC(ctx, out.tiles.ensure((size_t)T2 * kTileWords * 4));
        VDF_ALLOC(ctx, out.pc.ensure((size_t)T2 * kTile * 4));
        tc6_pack_kernel<kTile, false><<<T2, 128, 0, ctx->stream>>>(in, perm, n, out.tiles.as<uint32_t>(), out.pc.as<uint32_t>(), d_pads_or);
        VDF_LAUNCHED(ctx);
        return VDF_OK;
    }
    // cover n rounded up to whole 128-hash tiles (the tile ranges count those), in whole super-tiles: T2 * 96 is a
    // multiple of 192 and of 64
    const uint64_t n128 = (n + kTile - 1) / kTile * kTile;
    const uint32_t T = (uint32_t)((n128 + kT6Cols - 1) / kT6Cols), T2 = (T + 1) & ~1u;
    VDF_ALLOC(ctx, out.tiles.ensure((size_t)T2 * kT6ColTileBytes));
    VDF_ALLOC(ctx, out.pc.ensure((size_t)T2 * kT6Cols * 4));
    VDF_ALLOC(ctx, out.pcmin.ensure((size_t)T2 * kT6Cols / 64 * 4));
    tc6_pack_kernel<kT6Cols, true><<<T2, 128, 0, ctx->stream>>>(in, perm, n, out.tiles.as<uint32_t>(), out.pc.as<uint32_t>(), d_pads_or);
    VDF_LAUNCHED(ctx);
    const uint64_t groups = (uint64_t)T2 * kT6Cols / 64;
    pcmin64_kernel<<<(unsigned)((groups + 7) / 8), 256, 0, ctx->stream>>>(out.pc.as<uint32_t>(), groups, out.pcmin.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

static TcParams plan_params(const Plan& pl) {
    TcParams p;
    memset(&p, 0, sizeof p);
    p.row_lo = pl.row_lo.as<uint32_t>(), p.row_hi = pl.row_hi.as<uint32_t>();
    p.tile_range = pl.tile_range.as<uint2>();
    p.chunk = pl.chunk;
    p.rank = pl.rank, p.world = pl.world;
    p.unit_off = pl.unit_off.as<uint32_t>();
    p.n_pairs = (pl.n_row_tiles + 1) / 2;
    p.unit_list = pl.unit_list;
    p.peers.world = 0;
    return p;
}

// Work units of the CTA-pair kernels for the tile ranges in `pl` (no host round trip: the unit count lands in pl.stats[2],
// which the caller reads together with the other statistics of the plan).  The unit list is sized for its upper bound and
// padded with ~0 entries, which a stable sort on the chunk bits leaves behind the real ones.
int tc_plan_units(vdf_ctx* ctx, Plan& pl) {
    const uint32_t n_pairs = (pl.n_row_tiles + 1) / 2;
    uint32_t n_st, ch;
    if (pl.variant == 6) {
        n_st = (uint32_t)(((uint64_t)pl.n_col_tiles * kTile + 2 * kT6Cols - 1) / (2 * kT6Cols));
        // a unit costs ~5 us of set-up and drain (tensor-memory allocation, row operand load + expansion, cluster syncs,
        // the last epilogue) next to 0.84 us per super-tile: measured at 1 M hashes, 130.0 / 127.3 / 125.8 / 125.0 / 124.5 ms
        // for chunks of 128 / 256 / 512 / 1024 / 2048.  Long chunks, as long as >= 64 units per resident pair keep the tail short.
        ch = pl.tc_chunk ? pl.tc_chunk : 2048;
    } else {
        n_st = (pl.n_col_tiles + 1) / 2;
        ch = pl.tc_chunk ? pl.tc_chunk : 128;
    }
    while (!pl.tc_chunk && ch > 2 && (uint64_t)n_pairs * ((n_st + ch - 1) / ch) < (uint64_t)(ctx->sm_count / 2) * 64 * pl.world) ch >>= 1;
    while ((n_st + ch - 1) / ch > 65535) ch *= 2;
    pl.chunk = ch;
    VDF_ALLOC(ctx, pl.unit_cnt.ensure((size_t)(n_pairs + 1) * 4));
    VDF_ALLOC(ctx, pl.unit_off.ensure((size_t)(n_pairs + 1) * 4));
    TcParams p = plan_params(pl);
    if (pl.variant == 6) tc6_units_kernel<<<(n_pairs + 1 + 255) / 256, 256, 0, ctx->stream>>>(p, pl.n_row_tiles, pl.unit_cnt.as<uint32_t>());
    else tc5_units_kernel<<<(n_pairs + 1 + 255) / 256, 256, 0, ctx->stream>>>(p, pl.n_row_tiles, pl.unit_cnt.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    size_t tmp = 0;
    VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, pl.unit_cnt.as<uint32_t>(), pl.unit_off.as<uint32_t>(), (size_t)n_pairs + 1, ctx->stream));
    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
    VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tmp, pl.unit_cnt.as<uint32_t>(), pl.unit_off.as<uint32_t>(), (size_t)n_pairs + 1,
                                                ctx->stream));
    ctx->launches += 1;
    // stats[2] (u64, zeroed by the caller) <- the unit count (u32)
    VDF_CUDA(ctx, cudaMemcpyAsync(pl.stats.as<unsigned long long>() + 2, pl.unit_off.as<uint32_t>() + n_pairs, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    pl.unit_list = nullptr;
    if (pl.variant != 6) return VDF_OK;
    const uint32_t n_chunks = (n_st + ch - 1) / ch;
    const uint64_t ub = (uint64_t)n_pairs * ((n_chunks - 1) / pl.world + 1);  // a pair owns every world-th chunk of its range
    if (ub > 0x3FFFFFFFull) {
        ctx->err = "too many work units for one launch";
        return VDF_ERR_INVALID;
    }
    VDF_ALLOC(ctx, pl.unit_list_a.ensure((size_t)ub * 8));
    VDF_CUDA(ctx, cudaMemsetAsync(pl.unit_list_a.p, 0xFF, (size_t)ub * 8, ctx->stream));
    tc6_unit_list_kernel<<<(n_pairs + 255) / 256, 256, 0, ctx->stream>>>(p, pl.n_row_tiles, pl.unit_list_a.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    pl.unit_list = pl.unit_list_a.as<uint64_t>();
    if (pl.unit_order == 0 && n_chunks > 1) {  // chunk-major: stable radix sort on the chunk bits of the P-major list
        int cbits = 1;
        while ((1u << cbits) < n_chunks) ++cbits;
        VDF_ALLOC(ctx, pl.unit_list_b.ensure((size_t)ub * 8));
        size_t tmp2 = 0;
        VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp2, pl.unit_list_a.as<uint64_t>(), pl.unit_list_b.as<uint64_t>(), (size_t)ub, 32,
                                                     32 + cbits, ctx->stream));
        VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp2));
        VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp2, pl.unit_list_a.as<uint64_t>(), pl.unit_list_b.as<uint64_t>(), (size_t)ub,
                                                     32, 32 + cbits, ctx->stream));
        ctx->launches += 1;
        pl.unit_list = pl.unit_list_b.as<uint64_t>();
    }
    return VDF_OK;
}

int tc_launch(vdf_ctx* ctx, const Plan& pl, const Packed& rows, const Packed& cols, const uint32_t* row_id, uint64_t col_base, uint32_t tol,
              uint64_t capacity, unsigned long long* counter) {
    if (!ctx->tc_attrs) {  // per device, hence per context
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc5Smem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
        ctx->tc_attrs = true;
    }
    if (pl.n_units == 0) return VDF_OK;
    TcParams p = plan_params(pl);
    p.row_exp = rows.tiles.as<uint8_t>(), p.col_exp = cols.tiles.as<uint8_t>();
    p.row_pc = rows.pc.as<uint32_t>(), p.col_pc = cols.pc.as<uint32_t>(), p.col_pcmin = cols.pcmin.as<uint32_t>();
    p.row_id = row_id;
    p.keys = ctx->raw_keys.as<uint64_t>(), p.counter = counter, p.capacity = capacity, p.col_base = col_base;
    p.tol = tol > 1024u ? 1024u : tol;  // no distance exceeds 1024 (and the epilogue's threshold is a signed int)
    if (ctx->exchange) {
        if (pl.variant != 6 || !ctx->peer.world) {
            ctx->err = "the peer exchange needs search_variant 6 and vdf_peer_open";
            return VDF_ERR_INVALID;
        }
        p.peers = ctx->peer.ptrs((uint32_t)(ctx->peer.epoch & 1));
    }
    kt_begin(ctx, 0);
    if (pl.variant == 6) {
        const uint32_t n_col_st = (uint32_t)(((uint64_t)pl.n_col_tiles * kTile + 2 * kT6Cols - 1) / (2 * kT6Cols));
        const dim3 grid(2 * pl.n_units);
        if (ctx->tc_a_tmem && pl.fold) hamming_tc6_kernel<true, true><<<grid, kTc6Threads, kTc6Smem, ctx->stream>>>(p, pl.n_row_tiles, n_col_st);
        else if (ctx->tc_a_tmem) hamming_tc6_kernel<true, false><<<grid, kTc6Threads, kTc6Smem, ctx->stream>>>(p, pl.n_row_tiles, n_col_st);
        else if (pl.fold) hamming_tc6_kernel<false, true><<<grid, kTc6Threads, kTc6Smem, ctx->stream>>>(p, pl.n_row_tiles, n_col_st);
        else hamming_tc6_kernel<false, false><<<grid, kTc6Threads, kTc6Smem, ctx->stream>>>(p, pl.n_row_tiles, n_col_st);
    } else {
        hamming_tc5_kernel<<<2 * pl.n_units, kTc5Threads, kTc5Smem, ctx->stream>>>(p, pl.n_row_tiles);
    }
    kt_end(ctx, 0);
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

}  // namespace vdf
