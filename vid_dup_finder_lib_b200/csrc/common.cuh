// common.cuh -- context, scratch buffers and error plumbing shared by the kernels behind include/vdf_b200.h
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/vdf_b200.h"

namespace vdf {

constexpr int kTile = 128;                  // hashes per tile edge
constexpr int kWords32 = 32;                // 1024 bits as u32 words
constexpr int kTileWords = kTile * kWords32;  // 4096 u32 = 16 KB

// growable device buffer (never shrinks; owned by the context)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

// Edge exchange over NVLink peer memory (multi-GPU search, one process per GPU): every rank owns a buffer of two halves
// {arrived u64 @0, seg_count[8] u64 @128, keys[world][seg_cap] @256} that all peers map through CUDA IPC.  The pair kernel
// takes a slot from its LOCAL match counter and stores the key into segment `rank` of EVERY rank's buffer at that slot: plain
// remote stores over NVLink, fire and forget, overlapped with the math (a remote atomic per match was tried first: its round
// trip stalls the epilogue warp ~3 us per match and cost 1.4 ms per 62 ms launch on 2 GPUs).  At the end of the call each
// rank publishes its count into seg_count[rank] of every buffer and signals `arrived`; when all ranks have signalled, every
// rank holds the whole edge list: no collective call, no count exchange through the host.  Calls alternate between the
// halves, so a fast rank's next call never writes into memory a slow rank is still reading; `arrived` only grows.
constexpr int kMaxPeers = 8;
struct PeerPtrs {
    uint8_t* base[kMaxPeers];  // mapped buffer of every rank (base[rank] is the local one)
    uint64_t half_bytes;       // 256 + world * seg_cap * 8
    uint64_t seg_cap;          // keys one rank can contribute
    uint32_t world, half, rank;
    __host__ __device__ unsigned long long* arrived(uint32_t r) const {
        return reinterpret_cast<unsigned long long*>(base[r] + half * half_bytes);
    }
    __host__ __device__ unsigned long long* seg_count(uint32_t r, uint32_t writer) const {
        return reinterpret_cast<unsigned long long*>(base[r] + half * half_bytes + 128) + writer;
    }
    __host__ __device__ uint64_t* keys(uint32_t r, uint32_t writer) const {
        return reinterpret_cast<uint64_t*>(base[r] + half * half_bytes + 256) + writer * seg_cap;
    }
};
struct PeerExchange {
    uint32_t world = 0, rank = 0;  // world 0: not set up
    uint64_t cap = 0;              // keys per half over all writers
    void* local = nullptr;
    void* mapped[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t epoch = 0;  // exchange calls so far (all ranks make the same calls)
    PeerPtrs ptrs(uint32_t half) const {
        PeerPtrs pp;
        for (uint32_t r = 0; r < (uint32_t)kMaxPeers; ++r) pp.base[r] = static_cast<uint8_t*>(mapped[r]);
        pp.seg_cap = world ? cap / world : 0;
        pp.half_bytes = 256 + cap * 8, pp.world = world, pp.half = half, pp.rank = rank;
        return pp;
    }
};

// one axis of the Lanczos3 u8 resize (fast_image_resize Normalizer16) for a given input size, resident in HBM
struct CoefTable {
    uint32_t in_size = 0, window = 0, precision = 0;
    uint32_t* d_bounds = nullptr;  // 16 x (start, size)
    int16_t* d_k = nullptr;        // 16 x window
    int8_t* d_kb = nullptr;        // tensor-core operand: [2][16][in_size padded] split hi/lo bytes (built lazily)
    std::vector<uint32_t> h_bounds;
    std::vector<int16_t> h_k;
};

}  // namespace vdf

struct vdf_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};  // hash.cu: crops of chunk k are back on the host
    std::string err;
    uint32_t rank = 0, world = 1;
    uint64_t max_edges = 1ull << 28, initial_edges = 1ull << 22;
    int search_variant = 6;  // 0: plain XOR + POPC; 1: XOR + carry-save adder + POPC, 8x8 pairs/thread;
                             // 2: carry-save on 8x4 pairs/thread, two CTAs per SM; 3: tcgen05 kind::i8, byte-expanded
                             // tiles in HBM; 4: the same on CTA pairs; 5: CTA pairs, packed tiles, kind::i8;
                             // 6 (default): CTA pairs, packed tiles, kind::mxf4 on e2m1 {0, 1} operands (search_tc.cu)
    int hash_variant = 0;
    uint32_t tc_chunk = 0;  // column super-tiles per CTA-pair work unit (0: automatic)
    uint32_t tc_unit_order = 0;  // variant 6 work-unit order: 0 chunk-major (L2-friendly, default), 1 row-pair-major
    uint32_t tc_a_tmem = 1;     // variant 6: three quarters of the row operand in tensor memory (0: all of it in shared memory)
    uint32_t hash_chunks = 1;  // hash.cu: software-pipeline chunks per call (1: letterbox, then resize, over the whole batch)
    int exchange = 0;       // 1: searches append their matches to every rank's peer buffer (vdf_peer_*), see PeerExchange
    vdf::PeerExchange peer;
    int grouping = 0;       // 0: the reference's greedy rule (parity); 1: connected components (GPU union-find, group.cu)
    uint64_t launches = 0, h2d = 0, d2h = 0;
    double phase_ms[4] = {0, 0, 0, 0};  // last vdf_search*: host sort, gather + H2D enqueue, device, index remap (host.cu)
    // device time of the dominant kernels (CUDA events on `stream`): 0 hamming tiles, 1 resize, 2 letterbox, 3 dct+pack
    cudaEvent_t kt0[4] = {nullptr, nullptr, nullptr, nullptr}, kt1[4] = {nullptr, nullptr, nullptr, nullptr};
    bool kt_pending[4] = {false, false, false, false};
    double kt_ms[4] = {0, 0, 0, 0};
    uint64_t kt_n[4] = {0, 0, 0, 0};

    // search scratch
    vdf::DevBuf row_tiles, col_tiles, row_lo, row_hi, row_id, tile_range, raw_keys, sort_tmp, misc, keys_a, keys_b;
    vdf::DevBuf in_hash, in_dur, in_hash2, in_dur2, ref_perm, ref_key;
    vdf::DevBuf exp_rows, exp_cols, pc_rows, pc_cols, pcmin_rows, pcmin_cols, unit_cnt, unit_off, unit_list_a, unit_list_b;  // tensor-core search: byte-expanded tiles + popcounts
    // grouping scratch
    vdf::DevBuf g_rk, g_rks, g_state, g_parent, g_wl0, g_wla, g_wlb, g_mk, g_mks, g_flag, g_scan, g_gp, g_mem;
    // hashing scratch
    vdf::DevBuf h_frames[2], h_jobs, h_sides, h_crop, h_small, h_hash, h_desc;
    vdf::PinnedBuf pin_a, pin_b, pin_frames[2], h_groups;  // h_groups: staging of the group CSR on its way to the caller
    std::map<uint32_t, vdf::CoefTable> coef_cache;
    std::map<uint64_t, void*> bfrag_cache;  // (cropped width << 8 | shift) -> IMMA B fragments in HBM
    bool dct_consts_loaded = false;
};

#define VDF_CUDA(ctx, call)                                                                       \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                         std::to_string(__LINE__) + ")";                                          \
            return VDF_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define VDF_ALLOC(ctx, call)                                                                      \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                         std::to_string(__LINE__) + ")";                                          \
            cudaGetLastError();                                                                   \
            return VDF_ERR_ALLOC;                                                                 \
        }                                                                                         \
    } while (0)

#define VDF_LAUNCHED(ctx)                                                                         \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                         std::to_string(__LINE__) + ")";                                          \
            return VDF_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define VDF_TRY(expr)                \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != VDF_OK) return rc__; \
    } while (0)

namespace vdf {
// kernel timing helpers (api.cu)
void kt_begin(vdf_ctx* ctx, int which);
void kt_end(vdf_ctx* ctx, int which);
void kt_collect(vdf_ctx* ctx);
// search.cu
int search_self_device(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* d_dur, uint64_t n, uint32_t tol,
                       uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int search_refs_device(vdf_ctx* ctx, const uint64_t* d_cand, const uint32_t* d_cand_dur, uint64_t n_cand,
                       uint64_t cand_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                       uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int self_window_pairs(vdf_ctx* ctx, const uint32_t* d_dur, uint64_t n, uint64_t* pairs_out);
int sort_keys(vdf_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n);
// search_tc.cu
int tc_expand(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, DevBuf& exp, DevBuf& pc);
int tc_launch(vdf_ctx* ctx, uint32_t n_row_tiles, uint32_t n_col_tiles, uint32_t max_span_tiles, const uint8_t* row_exp,
              const uint8_t* col_exp, const uint32_t* row_pc, const uint32_t* col_pc, const uint32_t* col_pcmin,
              const uint32_t* row_id, uint64_t col_base, uint32_t tol, uint64_t capacity, unsigned long long* counter);
int tc5_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, DevBuf& tiles, DevBuf& pc, DevBuf& pcmin);
int tc6_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, bool as_columns, DevBuf& tiles, DevBuf& pc,
             DevBuf& pcmin);
// group.cu
int group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, vdf_groups* out);
int group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t n_edges, vdf_groups* out);
inline int group_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, vdf_groups* out) {
    return ctx->grouping == 1 ? group_components_device(ctx, n, d_keys_sorted, n_edges, out)
                              : group_greedy_device(ctx, n, d_keys_sorted, n_edges, out);
}
// hash.cu
int hash_stacks_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                       uint64_t* d_out_hash, uint8_t* d_out_small, int32_t* out_status, uint32_t* out_crop);
int hash_from_small_device(vdf_ctx* ctx, const uint8_t* d_small, uint32_t n, uint64_t* d_out_hash);
void free_coef_cache(vdf_ctx* ctx);
}  // namespace vdf
