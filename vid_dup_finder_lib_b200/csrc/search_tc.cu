// search_tc.cu -- tensor-core Hamming search, search_variant 3.
//
// hamming(a, b) = pc(a) + pc(b) - 2 * <a, b> with the 1024 bits of a hash expanded to 1024 bytes in {0, 1}: the
// all-pairs comparison becomes an exact u8 x u8 -> s32 contraction with K = 1024 that runs on the 5th-generation
// tensor cores (tcgen05.mma kind::i8, SASS UTCIMMA), accumulators in tensor memory.  A pair matches iff
//     2 * dot(i, j) - pc(j) >= pc(i) - tol.
// Results are bit-identical to the XOR+POPC kernels (integers throughout); this kernel is selected with
// search_variant = 3 and is validated against the same oracle tests.
//
// Layout in HBM (expand_tiles_kernel): exp[tile][kc 0..7][row 0..127][128 B], i.e. per 128-hash tile eight K-chunks of
// 128 bytes per row, each chunk block (16 KB) stored exactly as the UMMA K-major SWIZZLE_128B canonical layout wants it
// in shared memory (16-byte unit u of row r lives at unit u ^ (r & 7)), so plain 1-D TMA bulk copies stage operands.
//
// CTA = 6 warps, one per SM (192 KB of shared memory): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane,
// owns the TMEM allocation), warps 2-5 = epilogue (tcgen05.ld, one accumulator row per thread).  The row tile (A, 128
// hashes x 1024 B = 128 KB) stays resident; column super-tiles (B, 256 hashes) stream through a 2-stage ring of
// 32 KB K-chunks; each super-tile is 8 chunks x 4 MMAs of 128x256x32; two 256-column TMEM accumulators let the
// epilogue of super-tile t overlap the MMAs of t+1.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace vdf {

constexpr int kTcChunkBytes = kTile * 128;        // one K-chunk of one tile: 16 KB
constexpr int kTcTileBytes = 8 * kTcChunkBytes;   // expanded tile: 128 KB
constexpr int kTcStages = 2;
constexpr int kTcStageBytes = 2 * kTcChunkBytes;  // 256 columns x 128 B
constexpr int kTcThreads = 192;
constexpr size_t kTcSmem = (size_t)kTcTileBytes + kTcStages * kTcStageBytes + 1024 /*align*/ + 256 /*barriers*/;

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
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = unsigned 8 bit (0), both K-major,
// N = 256 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
constexpr uint32_t kTcIdesc = (2u << 4) | (32u << 17) | (8u << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kTcIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar))
                 : "memory");
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

// ------------------------------------------------------------------------------------------------ expansion
// bits -> {0,1} bytes in the swizzled K-major tile layout, plus per-hash popcounts.  grid = (tiles, 8 K-chunks).
__global__ void __launch_bounds__(256) expand_tiles_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm,
                                                           uint64_t n, uint8_t* __restrict__ exp, uint32_t* __restrict__ pc) {
    const uint64_t tile = blockIdx.x;
    const uint32_t kc = blockIdx.y;  // bits kc*128 .. kc*128+127 = u32 words kc*4 .. kc*4+3
    uint8_t* dst = exp + tile * (uint64_t)kTcTileBytes + (uint64_t)kc * kTcChunkBytes;
    for (uint32_t q = threadIdx.x; q < kTile * 8; q += 256) {
        const uint32_t row = q >> 3, unit = q & 7;  // 16-byte unit = 16 bits of the hash
        const uint64_t g = tile * kTile + row;
        uint32_t bits = 0;
        if (g < n) {
            const uint64_t src = perm ? perm[g] : g;
            bits = (in[src * 32 + kc * 4 + (unit >> 1)] >> ((unit & 1) * 16)) & 0xFFFFu;
        }
        uint4 v;  // nibble * 0x00204081 spreads bit i of the nibble to bit 0 of byte i
        v.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
        v.y = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.z = (((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
        v.w = (((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4*>(dst + row * 128 + ((unit ^ (row & 7)) << 4)) = v;
    }
    if (kc == 0) {
        for (uint32_t row = threadIdx.x; row < (uint32_t)kTile; row += 256) {
            const uint64_t g = tile * kTile + row;
            uint32_t c = 0;
            if (g < n) {
                const uint64_t src = perm ? perm[g] : g;
                for (int w = 0; w < 32; ++w) c += __popc(in[src * 32 + w]);
            }
            pc[g] = c;
        }
    }
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

__global__ void __launch_bounds__(kTcThreads, 1) hamming_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;
    uint8_t* sB = base + kTcTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kTcStages * kTcStageBytes);
    uint64_t* full = bars;                    // [kTcStages]
    uint64_t* empty = bars + kTcStages;       // [kTcStages]
    uint64_t* a_full = bars + 2 * kTcStages;  // [1]
    uint64_t* acc_full = a_full + 1;          // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const uint32_t I = blockIdx.x, c = blockIdx.y;
    if (p.world > 1 && ((I + c) % p.world) != p.rank) return;
    const uint2 rng = p.tile_range[I];
    const uint32_t st_begin = rng.x / 2, st_end = (rng.y + 1) / 2;  // super-tiles covering the 128-tile range
    const uint32_t st0 = st_begin + c * p.chunk;
    if (rng.x >= rng.y || st0 >= st_end) return;
    const uint32_t n_st = min(p.chunk, st_end - st0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) tc_mbar_init(&full[s], 1), tc_mbar_init(&empty[s], 1);
        tc_mbar_init(a_full, 1);
        for (int b = 0; b < 2; ++b) tc_mbar_init(&acc_full[b], 1), tc_mbar_init(&acc_empty[b], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the MMA warp owns all 512 TMEM columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer
            tc_mbar_expect_tx(a_full, kTcTileBytes);
            for (int kc = 0; kc < 8; ++kc)
                tc_bulk_g2s(sA + kc * kTcChunkBytes, p.row_exp + (size_t)I * kTcTileBytes + (size_t)kc * kTcChunkBytes,
                            kTcChunkBytes, a_full);
            uint32_t it = 0;
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint8_t* t0 = p.col_exp + (size_t)(2 * (st0 + s)) * kTcTileBytes;
                for (int kc = 0; kc < 8; ++kc, ++it) {
                    const uint32_t stage = it % kTcStages;
                    tc_mbar_wait(&empty[stage], ((it / kTcStages) & 1) ^ 1);
                    tc_mbar_expect_tx(&full[stage], kTcStageBytes);
                    uint8_t* d = sB + stage * kTcStageBytes;
                    tc_bulk_g2s(d, t0 + (size_t)kc * kTcChunkBytes, kTcChunkBytes, &full[stage]);
                    tc_bulk_g2s(d + kTcChunkBytes, t0 + kTcTileBytes + (size_t)kc * kTcChunkBytes, kTcChunkBytes, &full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer
            tc_mbar_wait(a_full, 0);
            const uint32_t a_addr = tc_smem_u32(sA), b_addr = tc_smem_u32(sB);
            uint32_t it = 0;
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t buf = s & 1;
                tc_mbar_wait(&acc_empty[buf], ((s >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kc = 0; kc < 8; ++kc, ++it) {
                    const uint32_t stage = it % kTcStages;
                    tc_mbar_wait(&full[stage], (it / kTcStages) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc_mma(d_tmem, tc_desc(a_addr + kc * kTcChunkBytes + ks * 32),
                               tc_desc(b_addr + stage * kTcStageBytes + ks * 32), (kc | ks) != 0);
                    tc_commit(&empty[stage]);  // the stage is free once these MMAs have read it
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else {  // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31, one accumulator row per thread
        const uint32_t quarter = warp & 3;
        const uint32_t row = quarter * 32 + lane;
        const uint32_t gi = I * kTile + row;
        const int thr = (int)p.row_pc[gi] - (int)p.tol;
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t buf = s & 1;
            tc_mbar_wait(&acc_full[buf], (s >> 1) & 1);
            tc_fence_after();
            const uint32_t col_first = (st0 + s) * 256;
            for (int q = 0; q < 4; ++q) {
                uint32_t v[64];
                __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the rare emit path
                tc_ld64(tmem_base + buf * 256 + q * 64 + ((quarter * 32) << 16), v);
                const uint4* pcj = reinterpret_cast<const uint4*>(p.col_pc + col_first + q * 64);
                int best = -0x7FFFFFFF;
#pragma unroll
                for (int k4 = 0; k4 < 16; ++k4) {
                    const uint4 pj = __ldg(pcj + k4);
                    const int t0 = 2 * (int)v[4 * k4 + 0] - (int)pj.x, t1 = 2 * (int)v[4 * k4 + 1] - (int)pj.y;
                    const int t2 = 2 * (int)v[4 * k4 + 2] - (int)pj.z, t3 = 2 * (int)v[4 * k4 + 3] - (int)pj.w;
                    v[4 * k4 + 0] = (uint32_t)t0, v[4 * k4 + 1] = (uint32_t)t1;
                    v[4 * k4 + 2] = (uint32_t)t2, v[4 * k4 + 3] = (uint32_t)t3;
                    best = max(best, max(max(t0, t1), max(t2, t3)));
                }
                if (best >= thr) {  // rare: at least one of these 64 pairs is within the tolerance
                    uint64_t mask = 0;
#pragma unroll
                    for (int k = 0; k < 64; ++k) mask |= (uint64_t)((int)v[k] >= thr) << k;
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
            if (lane == 0) tc_mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// search_variant 4: the same contraction on a CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256).  The pair
// owns a row super-tile of 256 hashes (each CTA keeps its own 128 rows resident) and each CTA stages only ITS half of
// every column super-tile (128 hashes x 128 B per K-chunk), so the L2 -> SM operand traffic and the shared-memory
// operand reads per pair of hashes are half of variant 3's, and the freed shared memory makes the ring 4 deep.
//   * full[stage] of the leader CTA counts two arrivals: its own producer's expect_tx and a relayed arrive from the
//     peer (an otherwise idle lane of the peer waits for its local bulk copies, then arrives on the leader's barrier
//     through a mapa'd shared::cluster address);
//   * the leader's tcgen05.commit multicasts to empty[stage] / acc_full[buf] of BOTH CTAs;
//   * acc_empty[buf] lives in the leader and counts the 8 epilogue warps of the pair.
// Column chunks are absolute (chunk c = super-tiles [c*chunk, (c+1)*chunk)), so that the pairs resident at the same
// time stream the same column tiles through L2.
constexpr int kTc2Stages = 4;
constexpr int kTc2StageBytes = kTcChunkBytes;  // this CTA's 128 columns x 128 B
constexpr size_t kTc2Smem = (size_t)kTcTileBytes + kTc2Stages * kTc2StageBytes + 1024 /*align*/ + 256 /*barriers*/;
// instruction descriptor as kTcIdesc with M = 256 (>> 4 at bit 24)
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
// the first capture of this kernel (profiles/r01_hamming_tc2_ncu.txt)
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1) hamming_tc2_kernel(const TcParams p, uint32_t n_row_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;
    uint8_t* sB = base + kTcTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kTc2Stages * kTc2StageBytes);
    uint64_t* full = bars;                     // [kTc2Stages]
    uint64_t* empty = bars + kTc2Stages;       // [kTc2Stages]
    uint64_t* a_full = bars + 2 * kTc2Stages;  // [1]
    uint64_t* acc_full = a_full + 1;           // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]  (leader's copy is the live one)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const uint32_t cr = tc_cluster_rank();  // 0 = leader (issues the MMAs), 1 = peer
    const uint32_t P = blockIdx.x >> 1, c = blockIdx.y;
    if (p.world > 1 && ((P + c) % p.world) != p.rank) return;
    // column super-tiles this pair needs: the union of its two row tiles' ranges
    const uint2 r0 = p.tile_range[2 * P];
    const uint2 r1 = (2 * P + 1 < n_row_tiles) ? p.tile_range[2 * P + 1] : make_uint2(0, 0);
    uint32_t t_lo = 0xFFFFFFFFu, t_hi = 0;
    if (r0.x < r0.y) t_lo = r0.x, t_hi = r0.y;
    if (r1.x < r1.y) t_lo = min(t_lo, r1.x), t_hi = max(t_hi, r1.y);
    if (t_lo >= t_hi) return;
    const uint32_t st0 = max(t_lo / 2, c * p.chunk), st1 = min((t_hi + 1) / 2, (c + 1) * p.chunk);
    if (st0 >= st1) return;  // the same decision in both CTAs of the pair
    const uint32_t n_st = st1 - st0;
    const uint32_t I = 2 * P + cr;  // this CTA's row tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        const uint32_t producers = cr == 0 ? 2 : 1;  // leader: own expect_tx + the peer's relayed arrive
        for (int s = 0; s < kTc2Stages; ++s) tc_mbar_init(&full[s], producers), tc_mbar_init(&empty[s], 1);
        tc_mbar_init(a_full, producers);
        for (int b = 0; b < 2; ++b) tc_mbar_init(&acc_full[b], 1), tc_mbar_init(&acc_empty[b], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates the pair's tensor memory (512 columns each)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();  // barriers of both CTAs are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer: this CTA's 128 rows, then its half of every column super-tile
            tc_mbar_expect_tx(a_full, kTcTileBytes);
            for (int kc = 0; kc < 8; ++kc)
                tc_bulk_g2s(sA + kc * kTcChunkBytes, p.row_exp + (size_t)I * kTcTileBytes + (size_t)kc * kTcChunkBytes,
                            kTcChunkBytes, a_full);
            uint32_t it = 0;
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint8_t* t0 = p.col_exp + (size_t)(2 * (st0 + s) + cr) * kTcTileBytes;
                for (int kc = 0; kc < 8; ++kc, ++it) {
                    const uint32_t stage = it % kTc2Stages;
                    tc_mbar_wait(&empty[stage], ((it / kTc2Stages) & 1) ^ 1);
                    tc_mbar_expect_tx(&full[stage], kTc2StageBytes);
                    tc_bulk_g2s(sB + stage * kTc2StageBytes, t0 + (size_t)kc * kTcChunkBytes, kTcChunkBytes, &full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        if (cr == 0) {  // ===== MMA issuer (leader CTA only): converged warp, one elected lane issues
            tc_mbar_wait(a_full, 0);
            const uint64_t a_desc = tc_desc(tc_smem_u32(sA)), b_desc = tc_desc(tc_smem_u32(sB));
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t buf = s & 1;
                tc_mbar_wait(&acc_empty[buf], ((s >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {  // it = 8 s + kc: stage and parity are compile-time (8 = 2 x kTc2Stages)
                    static_assert(kTc2Stages == 4, "stage/parity folding below assumes 4 stages");
                    tc_mbar_wait(&full[kc & 3], (kc >> 2) & 1);
                    tc_fence_after();
                    if (tc_elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc2_mma(d_tmem, a_desc + ((kc * kTcChunkBytes + ks * 32) >> 4),
                                    b_desc + (((kc & 3) * kTc2StageBytes + ks * 32) >> 4), (kc | ks) != 0);
                        tc2_commit(&empty[kc & 3]);  // frees the stage in both CTAs
                        if (kc == 7) tc2_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        } else if (lane == 0) {  // ===== peer CTA: relay "my half has landed" to the leader's barriers
            tc_mbar_wait(a_full, 0);
            tc_mbar_arrive_remote(a_full, 0);
            const uint32_t total = n_st * 8;
            for (uint32_t it = 0; it < total; ++it) {
                const uint32_t stage = it % kTc2Stages;
                tc_mbar_wait(&full[stage], (it / kTc2Stages) & 1);
                tc_mbar_arrive_remote(&full[stage], 0);
            }
        }
    } else {  // ===== epilogue (both CTAs): warp w reads TMEM lanes 32*(w%4) .. +31 of its own CTA
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
            for (int q = 0; q < 4; ++q) {
                uint32_t v[64];
                __syncwarp();
                tc_ld64(tmem_base + buf * 256 + q * 64 + ((quarter * 32) << 16), v);
                const uint4* pcj = reinterpret_cast<const uint4*>(p.col_pc + col_first + q * 64);
                int best = -0x7FFFFFFF;
#pragma unroll
                for (int k4 = 0; k4 < 16; ++k4) {
                    const uint4 pj = __ldg(pcj + k4);
                    const int t0 = 2 * (int)v[4 * k4 + 0] - (int)pj.x, t1 = 2 * (int)v[4 * k4 + 1] - (int)pj.y;
                    const int t2 = 2 * (int)v[4 * k4 + 2] - (int)pj.z, t3 = 2 * (int)v[4 * k4 + 3] - (int)pj.w;
                    v[4 * k4 + 0] = (uint32_t)t0, v[4 * k4 + 1] = (uint32_t)t1;
                    v[4 * k4 + 2] = (uint32_t)t2, v[4 * k4 + 3] = (uint32_t)t3;
                    best = max(best, max(max(t0, t1), max(t2, t3)));
                }
                if (best >= thr) {  // rare
                    uint64_t mask = 0;
#pragma unroll
                    for (int k = 0; k < 64; ++k) mask |= (uint64_t)((int)v[k] >= thr) << k;
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
    tc_cluster_sync();  // neither CTA may exit (or free tensor memory) while its partner can still touch it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ CTA pair, packed operands
// search_variant 5: variant 4 with the bit -> byte expansion moved INTO the kernel.  HBM keeps only packed tiles
// (pk[tile][K-chunk 0..7][hash 0..127][4 x u32], 16 KB per 128 hashes), so the operand traffic from L2 is 1/8 of variant
// 4's and no byte-expanded copy of the table (1 GB per million hashes) exists.  Four expander warps turn each K-chunk of
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

int tc5_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, DevBuf& tiles, DevBuf& pc, DevBuf& pcmin) {
    const uint32_t T = (uint32_t)((n + kTile - 1) / kTile);
    const uint32_t T2 = (T + 1) & ~1u;  // CTA pairs read tile pairs: one zero tile of padding when T is odd
    VDF_ALLOC(ctx, tiles.ensure((size_t)T2 * kTileWords * 4));
    VDF_ALLOC(ctx, pc.ensure((size_t)T2 * kTile * 4));
    VDF_ALLOC(ctx, pcmin.ensure((size_t)T2 * 2 * 4));
    tc5_pack_kernel<<<T2, kTile, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_hash), perm, n, tiles.as<uint32_t>(),
                                                  pc.as<uint32_t>(), pcmin.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

// ------------------------------------------------------------------------------------------------ CTA pair, 4-bit operands
// search_variant 6 (experimental): variant 5 with the bits expanded to e2m1 nibbles ({0, 1.0}) and
// tcgen05.mma kind::mxf4 (block-scaled, K = 64 per instruction, fp32 accumulators: exact, every partial sum is an
// integer <= 1024), which issues at twice the kind::i8 rate.  Block scales are the constant 1.0 (UE8M0 0x7F): the tensor-
// memory columns the scale factors live in are filled with 0x7F bytes once, so their layout does not matter.  Those
// columns have to come out of the accumulators' 512, hence super-tiles of 192 columns (two 192-column accumulators at
// [0,192) and [192,384), scale factors in [384,512)) and column tiles of 96 hashes.  Operands are half the size of
// variant 5's (A: 64 KB per CTA), which pays for an 8-stage ring.  K order: nibble j of output word m of a unit holds bit
// 4j + m of the packed word -- any permutation of K is fine as long as A and B use the same one.
constexpr int kT6Cols = 96;                          // hashes per column tile (one CTA's half of a super-tile)
constexpr int kT6Chunk = 128;                        // bytes per row per K-chunk = 256 bits
constexpr int kT6ABytes = 4 * kTile * kT6Chunk;      // 64 KB
constexpr int kT6StageBytes = kT6Cols * kT6Chunk;    // 12 KB
constexpr int kT6Stages = 8;
constexpr int kT6Expanders = 4;                      // expander warps (24 four-row groups per stage: 6 each)
constexpr int kTc6Threads = (6 + kT6Expanders) * 32; // warp 0 producer, 1 MMA, 2-5 epilogue, 6.. expanders
constexpr int kT6PackedBytes = kT6Cols * 128;        // packed column tile: 96 x 128 B = 12 KB; packed row tile: 16 KB
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
// one packed u32 -> one 16-byte unit of e2m1 nibbles (0x2 = 1.0)
__device__ __forceinline__ uint4 tc6_expand(uint32_t w) {
    return make_uint4((w << 1) & 0x22222222u, w & 0x22222222u, (w >> 1) & 0x22222222u, (w >> 2) & 0x22222222u);
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

// hashes [n][32] u32 -> pk[tile][K-chunk 0..3][row 0..T-1][8 x u32] + pc + pcmin64 (as tc5_pack_kernel, tile size T)
template <int T>
__global__ void __launch_bounds__(128) tc6_pack_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ perm, uint64_t n,
                                                       uint32_t* __restrict__ pk, uint32_t* __restrict__ pc) {
    if ((int)threadIdx.x >= T) return;
    const uint64_t g = (uint64_t)blockIdx.x * T + threadIdx.x;
    uint4* out = reinterpret_cast<uint4*>(pk + (size_t)blockIdx.x * T * 32) + threadIdx.x * 2;
    uint32_t c = 0;
    const uint4* src = g < n ? reinterpret_cast<const uint4*>(in) + (perm ? perm[g] : g) * 8 : nullptr;
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // q = 16-byte piece of the hash; K-chunk kc = q / 2 holds pieces 2kc, 2kc+1
        uint4 v = make_uint4(0, 0, 0, 0);
        if (src) v = src[q];
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        out[(q >> 1) * T * 2 + (q & 1)] = v;
    }
    pc[g] = c;
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

// kATmem: K-chunks 0-2 of the row operand live in tensor memory (columns [384, 480), written once per unit with tcgen05.st)
// and only chunk 3 in shared memory; scale factors in [480, 512).  The shared-memory data pipe is what bounds the all-smem
// form (ncu: 50 % tensor-core operand reads + 32 % expander stores at 86 % tensor-pipe activity); reading three quarters of
// A from tensor memory takes the operand reads from 7 KB to 4 KB per instruction.
template <bool kATmem>
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
                const uint4 e = tc6_expand(pw[w]);
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
            *reinterpret_cast<uint4*>(sA + (kATmem ? 0 : kc) * (kTile * kT6Chunk) + row * 128 + ((w ^ (row & 7)) << 4)) =
                tc6_expand(sP[(kc * kTile + row) * 8 + w]);
        }
    }
    tc_fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();
    tc_fence_after();
    const uint32_t sfa = tmem_base + kSfCol, sfb = tmem_base + kSfCol + (kATmem ? 16u : 64u);

    if (warp == 0) {
        if (lane == 0) {  // ===== bulk-copy producer: this CTA's packed column tile (96 hashes) of every super-tile
            for (uint32_t s = 0; s < n_st; ++s) {
                const uint32_t pb = s % kPB;
                tc_mbar_wait(&pempty[pb], ((s / kPB) & 1) ^ 1);
                tc_mbar_expect_tx(&pfull[pb], kT6PackedBytes);
                tc_bulk_g2s(sP + pb * kTileWords, col_tiles + (size_t)(2 * (st0 + s) + cr) * (kT6PackedBytes / 4), kT6PackedBytes,
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
            for (int kc = 0; kc < 4; ++kc) {
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
                    *reinterpret_cast<uint4*>(dst + row * 128 + ((w ^ (row & 7)) << 4)) = tc6_expand(bits[g]);
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
        const int thr = live ? (int)p.row_pc[gi] - (int)p.tol : 0x7FFFFFFF;
        const float thr_f = live ? (float)thr - 8388608.0f : 3.0e38f;  // 2 dot - pc(j) - 2^23 >= thr - 2^23, all exact in fp32
        for (uint32_t s = 0; s < n_st; ++s) {
            const uint32_t buf = s & 1;
            tc_mbar_wait(&acc_full[buf], (s >> 1) & 1);
            tc_fence_after();
            const uint32_t col_first = (st0 + s) * (2 * kT6Cols);
            const uint32_t* pcmin = p.col_pcmin + (col_first >> 6);
            for (int q = 0; q < 3; ++q) {
                uint32_t v[64];
                const int floor_pc = (int)__ldg(pcmin + q);
                __syncwarp();
                tc_ld64(tmem_base + buf * 192 + q * 64 + ((quarter * 32) << 16), v);
                uint32_t best = 0;  // accumulators are non-negative floats: their bit patterns order like the values
#pragma unroll
                for (int k = 0; k < 64; k += 2) best = max(best, max(v[k], v[k + 1]));
                if (2 * (int)__uint_as_float(best) - floor_pc >= thr) {
                    // The bound above is loose (largest dot and smallest popcount of 64 columns rarely belong to the same
                    // column): at tolerance 0.4 it lets most groups of random hashes through.  Second screen, exact and on
                    // the FMA pipe only: the largest 2 dot - pc(j) of the group.  0x4B000000 | pc is the float 2^23 + pc, so
                    // fma(dot, 2, -(2^23 + pc)) = 2 dot - pc - 2^23 exactly - no int<->float conversions (XU pipe, 16 per
                    // clock: 64 of them per thread cost as much as the group's MMAs and made the 1 M launch 2.5x slower
                    // at tolerance 0.4).
                    const uint4* pcj = reinterpret_cast<const uint4*>(p.col_pc + col_first + q * 64);
                    float top = -3.0e38f;
#pragma unroll
                    for (int k4 = 0; k4 < 16; ++k4) {
                        const uint4 pj = __ldg(pcj + k4);
                        top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 0]), 2.0f, -__uint_as_float(0x4B000000u | pj.x)));
                        top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 1]), 2.0f, -__uint_as_float(0x4B000000u | pj.y)));
                        top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 2]), 2.0f, -__uint_as_float(0x4B000000u | pj.z)));
                        top = fmaxf(top, fmaf(__uint_as_float(v[4 * k4 + 3]), 2.0f, -__uint_as_float(0x4B000000u | pj.w)));
                    }
                    uint64_t mask = 0;
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
                    while (mask) {
                        const int k = __ffsll((long long)mask) - 1;
                        mask &= mask - 1;
                        const uint32_t gj = col_first + q * 64 + k;
                        if (gj >= p.row_lo[gi] && gj < p.row_hi[gi]) {
                            const uint64_t rid = p.row_id ? p.row_id[gi] : gi;
                            const uint64_t key = (rid << 32) | (uint64_t)(gj + p.col_base);
                            if (p.peers.world) {  // fused exchange: this rank's segment of every rank's buffer, over NVLink
                                const unsigned long long slot = atomicAdd(p.counter, 1ull);
                                if (slot < p.peers.seg_cap)
                                    for (uint32_t r = 0; r < p.peers.world; ++r) p.peers.keys(r, p.peers.rank)[slot] = key;
                            } else {
                                const unsigned long long slot = atomicAdd(p.counter, 1ull);
                                if (slot < p.capacity) p.keys[slot] = key;
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

// rows: tiles of 128 hashes (+ pc); columns: tiles of 96 hashes padded to whole super-tiles (+ pc, pcmin)
int tc6_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, bool as_columns, DevBuf& tiles, DevBuf& pc,
             DevBuf& pcmin) {
    const uint32_t* in = reinterpret_cast<const uint32_t*>(d_hash);
    if (!as_columns) {
        const uint32_t T = (uint32_t)((n + kTile - 1) / kTile), T2 = (T + 1) & ~1u;
        VDF_ALLOC(ctx, tiles.ensure((size_t)T2 * kTileWords * 4));
        VDF_ALLOC(ctx, pc.ensure((size_t)T2 * kTile * 4));
        tc6_pack_kernel<kTile><<<T2, 128, 0, ctx->stream>>>(in, perm, n, tiles.as<uint32_t>(), pc.as<uint32_t>());
        VDF_LAUNCHED(ctx);
        return VDF_OK;
    }
    // cover n rounded up to whole 128-hash tiles (the tile ranges count those), in whole super-tiles: T2 * 96 is a
    // multiple of 192 and of 64
    const uint64_t n128 = (n + kTile - 1) / kTile * kTile;
    const uint32_t T = (uint32_t)((n128 + kT6Cols - 1) / kT6Cols), T2 = (T + 1) & ~1u;
    VDF_ALLOC(ctx, tiles.ensure((size_t)T2 * kT6PackedBytes));
    VDF_ALLOC(ctx, pc.ensure((size_t)T2 * kT6Cols * 4));
    VDF_ALLOC(ctx, pcmin.ensure((size_t)T2 * kT6Cols / 64 * 4));
    tc6_pack_kernel<kT6Cols><<<T2, 128, 0, ctx->stream>>>(in, perm, n, tiles.as<uint32_t>(), pc.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    const uint64_t groups = (uint64_t)T2 * kT6Cols / 64;
    pcmin64_kernel<<<(unsigned)((groups + 7) / 8), 256, 0, ctx->stream>>>(pc.as<uint32_t>(), groups, pcmin.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

// ------------------------------------------------------------------------------------------------ host side
int tc_expand(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, DevBuf& exp, DevBuf& pc) {
    const uint32_t T = (uint32_t)((n + kTile - 1) / kTile);
    const uint32_t T2 = (T + 1) & ~1u;  // super-tiles read tile pairs: keep an even number of (zero) tiles
    VDF_ALLOC(ctx, exp.ensure((size_t)T2 * kTcTileBytes));
    VDF_ALLOC(ctx, pc.ensure((size_t)T2 * kTile * 4));
    if (T2 > T) {
        VDF_CUDA(ctx, cudaMemsetAsync(exp.as<uint8_t>() + (size_t)T * kTcTileBytes, 0, kTcTileBytes, ctx->stream));
        VDF_CUDA(ctx, cudaMemsetAsync(pc.as<uint32_t>() + (size_t)T * kTile, 0, kTile * 4, ctx->stream));
    }
    expand_tiles_kernel<<<dim3(T, 8), 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d_hash), perm, n,
                                                            exp.as<uint8_t>(), pc.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

int tc_launch(vdf_ctx* ctx, uint32_t n_row_tiles, uint32_t n_col_tiles, uint32_t max_span_tiles, const uint8_t* row_exp,
              const uint8_t* col_exp, const uint32_t* row_pc, const uint32_t* col_pc, const uint32_t* col_pcmin,
              const uint32_t* row_id, uint64_t col_base, uint32_t tol, uint64_t capacity, unsigned long long* counter) {
    static bool attr_done = false;
    if (!attr_done) {
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc2Smem));
        VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc5Smem));
        attr_done = true;
    }
    TcParams p;
    p.row_exp = row_exp, p.col_exp = col_exp, p.row_pc = row_pc, p.col_pc = col_pc, p.col_pcmin = col_pcmin;
    p.row_lo = ctx->row_lo.as<uint32_t>(), p.row_hi = ctx->row_hi.as<uint32_t>(), p.row_id = row_id;
    p.tile_range = ctx->tile_range.as<uint2>();
    p.keys = ctx->raw_keys.as<uint64_t>(), p.counter = counter, p.capacity = capacity, p.col_base = col_base;
    p.tol = tol > 1024u ? 1024u : tol;  // no distance exceeds 1024 (and the epilogue's threshold is a signed int)
    p.rank = ctx->rank, p.world = ctx->world;
    p.unit_off = nullptr, p.n_pairs = 0, p.unit_list = nullptr;
    p.peers.world = 0;
    if (ctx->exchange) {
        if (ctx->search_variant != 6 || !ctx->peer.world) {
            ctx->err = "the peer exchange needs search_variant 6 and vdf_peer_open";
            return VDF_ERR_INVALID;
        }
        p.peers = ctx->peer.ptrs((uint32_t)(ctx->peer.epoch & 1));
    }
    if (ctx->search_variant >= 4) {  // CTA pairs: 256-row super-tiles x absolute chunks of column super-tiles
        const uint32_t n_pairs = (n_row_tiles + 1) / 2, n_st = (n_col_tiles + 1) / 2;
        // long chunks amortise the per-unit set-up (operand A, tensor-memory allocation, cluster syncs); keep >= 64
        // units per resident pair for balance
        uint32_t chunk = ctx->tc_chunk ? ctx->tc_chunk : 128;
        while (!ctx->tc_chunk && chunk > 2 &&
               (uint64_t)n_pairs * ((n_st + chunk - 1) / chunk) < (uint64_t)(ctx->sm_count / 2) * 64 * ctx->world)
            chunk >>= 1;
        while ((n_st + chunk - 1) / chunk > 65535) chunk *= 2;
        p.chunk = chunk;
        if (ctx->search_variant == 6) {  // as variant 5, super-tiles of 192 columns
            static bool attr6 = false;
            if (!attr6) {
                VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
                VDF_CUDA(ctx, cudaFuncSetAttribute(hamming_tc6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTc6Smem));
                attr6 = true;
            }
            const uint32_t n_col_st = (uint32_t)(((uint64_t)n_col_tiles * kTile + 2 * kT6Cols - 1) / (2 * kT6Cols));
            // a unit costs ~5 us of set-up and drain (tensor-memory allocation, row operand load + expansion, cluster syncs,
            // the last epilogue) next to 0.84 us per super-tile: measured at 1 M hashes, 130.0 / 127.3 / 125.8 / 125.0 / 124.5 ms
            // for chunks of 128 / 256 / 512 / 1024 / 2048.  Long chunks, as long as >= 64 units per resident pair keep the tail short.
            uint32_t ch = ctx->tc_chunk ? ctx->tc_chunk : 2048;
            while (!ctx->tc_chunk && ch > 2 && (uint64_t)n_pairs * ((n_col_st + ch - 1) / ch) < (uint64_t)(ctx->sm_count / 2) * 64 * ctx->world)
                ch >>= 1;
            p.chunk = ch;
            p.n_pairs = n_pairs;
            VDF_ALLOC(ctx, ctx->unit_cnt.ensure((size_t)(n_pairs + 1) * 4));
            VDF_ALLOC(ctx, ctx->unit_off.ensure((size_t)(n_pairs + 1) * 4));
            p.unit_off = ctx->unit_off.as<uint32_t>();
            tc6_units_kernel<<<(n_pairs + 1 + 255) / 256, 256, 0, ctx->stream>>>(p, n_row_tiles, ctx->unit_cnt.as<uint32_t>());
            VDF_LAUNCHED(ctx);
            size_t tmp = 0;
            VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->unit_cnt.as<uint32_t>(), ctx->unit_off.as<uint32_t>(),
                                                        (size_t)n_pairs + 1, ctx->stream));
            VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
            VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tmp, ctx->unit_cnt.as<uint32_t>(),
                                                        ctx->unit_off.as<uint32_t>(), (size_t)n_pairs + 1, ctx->stream));
            uint32_t n_units = 0;
            VDF_CUDA(ctx, cudaMemcpyAsync(&n_units, ctx->unit_off.as<uint32_t>() + n_pairs, 4, cudaMemcpyDeviceToHost, ctx->stream));
            VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (n_units == 0) return VDF_OK;
            if (n_units > 0x3FFFFFFFu) {
                ctx->err = "too many work units for one launch";
                return VDF_ERR_INVALID;
            }
            {   // explicit unit list, chunk-major (stable radix sort on the chunk bits of a P-major list)
                VDF_ALLOC(ctx, ctx->unit_list_a.ensure((size_t)n_units * 8));
                VDF_ALLOC(ctx, ctx->unit_list_b.ensure((size_t)n_units * 8));
                tc6_unit_list_kernel<<<(n_pairs + 255) / 256, 256, 0, ctx->stream>>>(p, n_row_tiles, ctx->unit_list_a.as<uint64_t>());
                VDF_LAUNCHED(ctx);
                p.unit_list = ctx->unit_list_a.as<uint64_t>();
                const uint32_t n_chunks = (n_col_st + ch - 1) / ch;
                int cbits = 1;
                while ((1u << cbits) < n_chunks) ++cbits;
                if (ctx->tc_unit_order == 0 && n_chunks > 1) {
                    size_t tmp2 = 0;
                    VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp2, ctx->unit_list_a.as<uint64_t>(), ctx->unit_list_b.as<uint64_t>(),
                                                                 (size_t)n_units, 32, 32 + cbits, ctx->stream));
                    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp2));
                    VDF_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tmp2, ctx->unit_list_a.as<uint64_t>(),
                                                                 ctx->unit_list_b.as<uint64_t>(), (size_t)n_units, 32, 32 + cbits, ctx->stream));
                    ctx->launches += 1;
                    p.unit_list = ctx->unit_list_b.as<uint64_t>();
                }
            }
            kt_begin(ctx, 0);
            if (ctx->tc_a_tmem) hamming_tc6_kernel<true><<<2 * n_units, kTc6Threads, kTc6Smem, ctx->stream>>>(p, n_row_tiles, n_col_st);
            else hamming_tc6_kernel<false><<<2 * n_units, kTc6Threads, kTc6Smem, ctx->stream>>>(p, n_row_tiles, n_col_st);
            kt_end(ctx, 0);
            VDF_LAUNCHED(ctx);
            return VDF_OK;
        }
        if (ctx->search_variant == 5) {  // 1-D grid over exactly the units this rank owns
            p.n_pairs = n_pairs;
            VDF_ALLOC(ctx, ctx->unit_cnt.ensure((size_t)(n_pairs + 1) * 4));
            VDF_ALLOC(ctx, ctx->unit_off.ensure((size_t)(n_pairs + 1) * 4));
            p.unit_off = ctx->unit_off.as<uint32_t>();
            tc5_units_kernel<<<(n_pairs + 1 + 255) / 256, 256, 0, ctx->stream>>>(p, n_row_tiles, ctx->unit_cnt.as<uint32_t>());
            VDF_LAUNCHED(ctx);
            size_t tmp = 0;
            VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->unit_cnt.as<uint32_t>(), ctx->unit_off.as<uint32_t>(),
                                                        (size_t)n_pairs + 1, ctx->stream));
            VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
            VDF_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tmp, ctx->unit_cnt.as<uint32_t>(),
                                                        ctx->unit_off.as<uint32_t>(), (size_t)n_pairs + 1, ctx->stream));
            ctx->launches += 1;
            uint32_t n_units = 0;
            VDF_CUDA(ctx, cudaMemcpyAsync(&n_units, ctx->unit_off.as<uint32_t>() + n_pairs, 4, cudaMemcpyDeviceToHost, ctx->stream));
            VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (n_units == 0) return VDF_OK;
            if (n_units > 0x3FFFFFFFu) {
                ctx->err = "too many work units for one launch";
                return VDF_ERR_INVALID;
            }
            kt_begin(ctx, 0);
            hamming_tc5_kernel<<<2 * n_units, kTc5Threads, kTc5Smem, ctx->stream>>>(p, n_row_tiles);
            kt_end(ctx, 0);
            VDF_LAUNCHED(ctx);
            return VDF_OK;
        }
        dim3 grid(2 * n_pairs, (n_st + chunk - 1) / chunk);
        kt_begin(ctx, 0);
        hamming_tc2_kernel<<<grid, kTcThreads, kTc2Smem, ctx->stream>>>(p, n_row_tiles);
        kt_end(ctx, 0);
        VDF_LAUNCHED(ctx);
        return VDF_OK;
    }
    const uint32_t span_st = max_span_tiles / 2 + 2;  // super-tiles a row tile's range can touch
    uint32_t chunk = 32;
    while (chunk > 2 && (uint64_t)n_row_tiles * ((span_st + chunk - 1) / chunk) < (uint64_t)ctx->sm_count * 4 * ctx->world)
        chunk >>= 1;
    p.chunk = chunk;
    dim3 grid(n_row_tiles, (span_st + chunk - 1) / chunk);
    kt_begin(ctx, 0);
    hamming_tc_kernel<<<grid, kTcThreads, kTcSmem, ctx->stream>>>(p);
    kt_end(ctx, 0);
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

}  // namespace vdf
