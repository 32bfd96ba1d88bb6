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
    bool ipc = false;    // peers mapped through CUDA IPC (other processes); false: plain peer pointers of this process (mgpu.cu)
    PeerPtrs ptrs(uint32_t half) const {
        PeerPtrs pp;
        for (uint32_t r = 0; r < (uint32_t)kMaxPeers; ++r) pp.base[r] = static_cast<uint8_t*>(mapped[r]);
        pp.seg_cap = world ? cap / world : 0;
        pp.half_bytes = 256 + cap * 8, pp.world = world, pp.half = half, pp.rank = rank;
        return pp;
    }
};

// one frame stack as the hashing kernels see it (hash.cu, motion.cu)
struct StackDev {
    uint64_t offset, frame_stride;
    uint32_t width, height, pitch;
    int32_t status;
    uint32_t aligned;  // base, frame stride and pitch are multiples of 16: the tensor-core resize may take it
    uint32_t pad;
};

// one axis of the Lanczos3 u8 resize (fast_image_resize Normalizer16) for a given input size, resident in HBM
struct CoefTable {
    uint32_t in_size = 0, window = 0, precision = 0;
    uint32_t* d_bounds = nullptr;  // 16 x (start, size)
    int16_t* d_k = nullptr;        // 16 x window
    int8_t* d_kb = nullptr;        // tensor-core operand: [2][16][in_size padded] split hi/lo bytes (built lazily)
    void* d_kva = nullptr;         // the table as IMMA A fragments of the fused kernel's vertical pass (hash.cu get_table)
    std::vector<uint32_t> h_bounds;
    std::vector<int16_t> h_k;
};


// One side of the pair matrix in the form the pair kernels read: rows (tiles of 128 hashes) or columns (variant 6: tiles of
// 96 hashes with a per-column fold unit appended to every tile).  Built once per table by the pack kernels.
struct Packed {
    uint64_t n = 0;
    int variant = -1;
    bool as_columns = false;
    DevBuf tiles, pc, pcmin;
};

// Which (row pair, column chunk) work units this rank evaluates for a given set of row windows, and the launch geometry:
// everything between "the windows are known" and "launch the pair kernel".  Cached with the table for search_self.
struct Plan {
    DevBuf row_lo, row_hi, tile_range, unit_cnt, unit_off, unit_list_a, unit_list_b, stats;
    const uint64_t* unit_list = nullptr;
    uint32_t n_row_tiles = 0, n_col_tiles = 0, n_units = 0, chunk = 0, max_span = 0;
    uint64_t pairs = 0;
    bool fold = false;  // both operands have zero pad bits: the in-contraction screen applies (search_tc.cu)
    // what the plan was made for
    bool valid = false;
    int variant = -1;
    uint32_t rank = 0, world = 1, tc_chunk = 0, unit_order = 0;
    int tc_fold = -1;
    void release() {
        DevBuf* b[] = {&row_lo, &row_hi, &tile_range, &unit_cnt, &unit_off, &unit_list_a, &unit_list_b, &stats};
        for (DevBuf* x : b) x->release();
        valid = false;
    }
};

// `Search::from(hashes)` (search_algorithm.rs:188-198) with the sorted table resident in HBM and prepared for the kernels
struct Table {
    uint64_t n = 0;
    const uint64_t* d_hash = nullptr;  // [n][16] u64; sorted entry k is row d_perm[k] when d_perm is given, else row k
    const uint32_t* d_perm = nullptr;
    const uint32_t* d_dur = nullptr;   // [n] durations in SORTED order
    Packed rows, cols;                 // row role / column role (variant 6 packs them differently; the others share `rows`)
    Plan self_plan;
    DevBuf meta;                       // [0] u32: OR over all hashes of the pad bits 1000..1023 (no real VideoHash sets them)
    uint32_t pads = 0;
    bool pads_known = false;
    void release() {
        DevBuf* b[] = {&rows.tiles, &rows.pc, &rows.pcmin, &cols.tiles, &cols.pc, &cols.pcmin, &meta};
        for (DevBuf* x : b) x->release();
        self_plan.release();
        rows.n = cols.n = 0;
        rows.variant = cols.variant = -1;
    }
};

}  // namespace vdf

struct vdf_table {
    vdf_ctx* ctx = nullptr;
    vdf::Table t;
    vdf::DevBuf own_hash, own_dur;  // when the table was created from host memory
};

struct vdf_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t lb_stream = nullptr;  // hash.cu: the letterbox scan of chunk k+1 runs beside the resize of chunk k
    cudaEvent_t ev_in = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};  // hash.cu: crops of chunk k are back on the host
    std::string err;
    uint32_t rank = 0, world = 1;
    uint64_t max_edges = 1ull << 28, initial_edges = 1ull << 22;
    int search_variant = 6;  // 0: plain XOR + POPC; 1: XOR + carry-save adder + POPC, 8x8 pairs/thread;
                             // 2: carry-save on 8x4 pairs/thread, two CTAs per SM; 5: tcgen05 kind::i8 on CTA pairs, packed
                             // tiles; 6 (default): CTA pairs, packed tiles, kind::mxf4 on e2m1 operands (search_tc.cu).
                             // (3 and 4, the byte-expanded steps towards 5, were removed in round 2: strictly dominated)
    int hash_variant = 0;
    uint32_t tc_chunk = 0;  // column super-tiles per CTA-pair work unit (0: automatic)
    uint32_t tc_unit_order = 0;  // variant 6 work-unit order: 0 chunk-major (L2-friendly, default), 1 row-pair-major
    uint32_t tc_a_tmem = 1;     // variant 6: three quarters of the row operand in tensor memory (0: all of it in shared memory)
    int tc_fold = -1;           // variant 6: -1 fold C - pc(j) into the contraction whenever both operands have zero pad bits
                                // (every real VideoHash), 0 never (the popcount-screen epilogue of round 1)
    uint64_t peer_timeout_ms = 0;  // peer exchange barrier: 0 = automatic (20 s + 1 ms per 2^26 pairs in the windows)
    // hash.cu: the letterbox scan may run on a second stream, chunk k+1 beside the resize of chunk k.  Measured (256 1080p stacks):
    // 1.88 ms per call overlapped in 4 chunks against 1.59 serial -- the latency-bound scan CTAs hold shared memory that a resize
    // CTA needs, and with HBM saturated their own loads crawl (1.08 ms on their stream against 0.19 alone); chunks also add
    // three kernel tails.  Default: one chunk, one stream, and a scan that is cheap on its own.
    uint32_t hash_chunks = 1;
    uint32_t hash_overlap = 0;
    uint32_t hash_fused = 1;     // 1 (default): hash_fused_kernel, one persistent launch per call does letterbox -> ... -> pack; 0: the
                                 // per-frame kernels of round 1 / early round 2 (letterbox scan kernels, resize_mma_kernel)
    uint32_t hash_fuse_dct = 1;  // DCT + threshold + pack in the resize kernel (the CTA that finishes a stack); 0: a kernel of its own
    int exchange = 0;       // 1: searches append their matches to every rank's peer buffer (vdf_peer_*), see PeerExchange
    vdf::PeerExchange peer;
    bool peer_dead = false;  // an exchange call failed mid-way: counters of the ranks disagree, vdf_peer_close / alloc / open again
    int grouping = 0;       // 0: the reference's greedy rule (parity); 1: connected components (GPU union-find, group.cu)
    // a context made by vdf_ctx_create_multi drives several GPUs of one node from one process: sub[0] is the context itself,
    // sub[1..] are contexts on the other devices (owned); host threads, one per device, run the shards (mgpu.cu)
    int sub_count = 0;
    vdf_ctx* sub[vdf::kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t launches = 0, h2d = 0, d2h = 0;
    double phase_ms[4] = {0, 0, 0, 0};  // last vdf_search*: host sort, gather + H2D enqueue, device, index remap (host.cu)
    // device time of the dominant kernels (CUDA events on `stream`): 0 hamming tiles, 1 resize, 2 letterbox, 3 dct+pack
    cudaEvent_t kt0[4] = {nullptr, nullptr, nullptr, nullptr}, kt1[4] = {nullptr, nullptr, nullptr, nullptr};
    bool kt_pending[4] = {false, false, false, false};
    double kt_ms[4] = {0, 0, 0, 0};
    uint64_t kt_n[4] = {0, 0, 0, 0};

    // search scratch: tables built per call by the entry points that take raw arrays (a vdf_table keeps its own)
    vdf::Table tmp_self, tmp_cand;
    vdf::Packed ref_rows;  // the references of one vdf_search_refs* call, in duration order
    vdf::Plan ref_plan;
    vdf::DevBuf raw_keys, sort_tmp, misc, keys_a, keys_b;
    vdf::DevBuf in_hash, in_dur, in_hash2, in_dur2, ref_perm, ref_key;
    vdf::DevBuf sk_a, sk_b, sk_c, sk_d, sk_order, sk_rank;  // host.cu: (duration, path prefix) keys of the GPU-side sort
    // grouping scratch
    vdf::DevBuf g_rk, g_rks, g_state, g_parent, g_wl0, g_wla, g_wlb, g_mk, g_mks, g_flag, g_scan, g_gp, g_mem, g_cnt, g_gstart;
    int greedy_blocks_per_sm = 0;
    // cudaFuncSetAttribute is per DEVICE: every context opts its kernels in once (a process may hold contexts on several GPUs)
    bool ham_attrs = false, tc_attrs = false;
    size_t hash_smem_set[3] = {0, 0, 0};
    // hashing scratch
    vdf::DevBuf h_frames[2], h_jobs, h_sides, h_crop, h_small, h_hash, h_desc, h_coef_lut, h_bfrag_lut, h_done, h_lbwork, h_fctl;
    vdf::DevBuf m_state, m_lut, m_sides, m_acc, m_a, m_b, m_c, m_f32, m_label;  // motion.cu scratch
    vdf::PinnedBuf pin_a, pin_b, pin_c, pin_frames[2], h_groups;  // h_groups: staging of the group CSR on its way to the caller
    vdf::PinnedBuf h_misc;  // landing zone of the few counters the host reads back per call
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
            cudaGetLastError();                                                                   \
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
void kt_begin(vdf_ctx* ctx, int which, cudaStream_t stream = nullptr);  // nullptr: the context's main stream
void kt_end(vdf_ctx* ctx, int which, cudaStream_t stream = nullptr);
void kt_collect(vdf_ctx* ctx);
// search.cu
int table_prepare(vdf_ctx* ctx, Table& t, const uint64_t* d_hash, const uint32_t* d_perm, const uint32_t* d_dur_sorted, uint64_t n,
                  bool for_self, bool for_cand);
int table_search_self(vdf_ctx* ctx, Table& t, uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int table_search_refs(vdf_ctx* ctx, Table& cand, uint64_t cand_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                      uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int search_self_device(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* d_dur, uint64_t n, uint32_t tol,
                       uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int search_refs_device(vdf_ctx* ctx, const uint64_t* d_cand, const uint32_t* d_cand_dur, uint64_t n_cand,
                       uint64_t cand_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                       uint32_t tol, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int self_window_pairs(vdf_ctx* ctx, const uint32_t* d_dur, uint64_t n, uint64_t* pairs_out);
int sort_keys(vdf_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, int end_bit = 64);
int bits_for(uint64_t n);
// search_tc.cu
int tc5_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, Packed& out);
int tc6_pack(vdf_ctx* ctx, const uint64_t* d_hash, const uint32_t* perm, uint64_t n, bool as_columns, Packed& out, uint32_t* d_pads_or);
int tc_plan_units(vdf_ctx* ctx, Plan& plan);
int tc_launch(vdf_ctx* ctx, const Plan& plan, const Packed& rows, const Packed& cols, const uint32_t* row_id, uint64_t col_base,
              uint32_t tol, uint64_t capacity, unsigned long long* counter);
// host.cu
int stage_and_sort(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* dur, const char* paths, const uint64_t* off, uint64_t n,
                   DevBuf& d_hash, PinnedBuf& pin_h, double* t_keys, double* t_stage);
uint32_t tolerance_to_int(double tolerance);
int ref_keys_to_csr(vdf_ctx* ctx, const uint64_t* d_keys, uint64_t ne, const uint32_t* d_order, uint64_t n_ref, vdf_csr* out);
// api.cu
int hash_stacks_host(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect, uint64_t* out_hash,
                     int32_t* out_status, uint32_t* out_crop);
void peer_release(vdf_ctx* ctx);
// mgpu.cu
int mgpu_search(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n,
                double tolerance, vdf_groups* out);
int mgpu_search_refs(vdf_ctx* ctx, const uint64_t* ref_hashes, const uint32_t* ref_durations, uint64_t n_ref, const uint64_t* cand_hashes,
                     const uint32_t* cand_durations, const char* cand_path_blob, const uint64_t* cand_path_off, uint64_t n_cand,
                     double tolerance, vdf_csr* out);
int mgpu_hash_stacks(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect, uint64_t* out_hash,
                     int32_t* out_status, uint32_t* out_crop);
// group.cu: host memory of results handed to the caller (a pool of pinned buffers; plain malloc when the pool is lent out)
void* result_alloc(size_t bytes);
void result_free(void* p);
// group.cu (d_remap, optional: sorted position -> the caller's index, applied to every member that is written out)
int group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, const uint32_t* d_remap, vdf_groups* out);
int group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t n_edges, const uint32_t* d_remap, vdf_groups* out);
inline int group_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, const uint32_t* d_remap, vdf_groups* out) {
    return ctx->grouping == 1 ? group_components_device(ctx, n, d_keys_sorted, n_edges, d_remap, out)
                              : group_greedy_device(ctx, n, d_keys_sorted, n_edges, d_remap, out);
}
// hash.cu
int hash_stacks_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                       uint64_t* d_out_hash, uint8_t* d_out_small, int32_t* out_status, uint32_t* out_crop);
int hash_from_small_device(vdf_ctx* ctx, const uint8_t* d_small, uint32_t n, uint64_t* d_out_hash);
void free_coef_cache(vdf_ctx* ctx);
int letterbox_all_frames(vdf_ctx* ctx, const uint8_t* d_frames, const StackDev* d_sd, uint32_t n, const uint8_t* d_luts, uint32_t* d_sides,
                         uint32_t* d_crop);
// motion.cu: Cropdetect::Motion (autocrop_frames.rs:36-218) for n stacks -> d_crop [n][4] (left, right, top, bottom)
int motion_crop_device(vdf_ctx* ctx, const uint8_t* d_frames, const StackDev* d_sd, const StackDev* h_sd, uint32_t n, uint32_t* d_crop);
}  // namespace vdf
