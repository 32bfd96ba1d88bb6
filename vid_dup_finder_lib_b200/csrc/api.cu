// api.cu -- the extern "C" boundary declared in include/vdf_b200.h: context management, host <-> HBM staging,
// and thin wrappers over the device-resident implementations in search.cu / group.cu / hash.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace vdf;

extern "C" {

const char* vdf_version(void) { return "vdf_b200 0.2.0 (sm_100a)"; }

int vdf_ctx_create(int device_id, vdf_ctx** out) {
    if (!out) return VDF_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_id < 0 || device_id >= count) {
        cudaGetLastError();
        return VDF_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return VDF_ERR_NO_DEVICE;
    if (prop.major != 10) return VDF_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
    if (cudaSetDevice(device_id) != cudaSuccess) return VDF_ERR_NO_DEVICE;
    vdf_ctx* ctx = new (std::nothrow) vdf_ctx();
    if (!ctx) return VDF_ERR_ALLOC;
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return VDF_ERR_CUDA;
    }
    for (int k = 0; k < 2; ++k) {
        cudaEventCreateWithFlags(&ctx->ev_copy[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_chunk[2 * k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_chunk[2 * k + 1], cudaEventDisableTiming);
    }
    for (int k = 0; k < 4; ++k) {
        cudaEventCreate(&ctx->kt0[k]);
        cudaEventCreate(&ctx->kt1[k]);
    }
    *out = ctx;
    return VDF_OK;
}


void vdf_ctx_destroy(vdf_ctx* ctx) {
    if (!ctx) return;
    for (int r = 1; r < ctx->sub_count; ++r) {  // a multi-GPU context owns the contexts of its other devices
        if (ctx->sub[r]) vdf_ctx_destroy(ctx->sub[r]);
        ctx->sub[r] = nullptr;
    }
    ctx->sub_count = 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->lb_stream) cudaStreamSynchronize(ctx->lb_stream);
    DevBuf* bufs[] = {&ctx->raw_keys,  &ctx->sort_tmp,  &ctx->misc,     &ctx->keys_a,  &ctx->keys_b,  &ctx->in_hash,
                      &ctx->in_dur,    &ctx->in_hash2,  &ctx->in_dur2,  &ctx->ref_perm, &ctx->ref_key, &ctx->g_rk,
                      &ctx->g_rks,     &ctx->g_state,   &ctx->g_parent, &ctx->g_wl0,   &ctx->g_wla,   &ctx->g_wlb,
                      &ctx->g_mk,      &ctx->g_mks,     &ctx->g_flag,   &ctx->g_scan,  &ctx->g_gp,    &ctx->g_mem,
                      &ctx->g_cnt,     &ctx->g_gstart,  &ctx->sk_a,     &ctx->sk_b,    &ctx->sk_c,    &ctx->sk_d,
                      &ctx->sk_order,  &ctx->sk_rank,   &ctx->ref_rows.tiles, &ctx->ref_rows.pc, &ctx->ref_rows.pcmin,
                      &ctx->h_frames[0], &ctx->h_frames[1], &ctx->h_jobs, &ctx->h_sides, &ctx->h_crop, &ctx->h_small,
                      &ctx->h_hash,    &ctx->h_desc,    &ctx->h_done,   &ctx->h_lbwork, &ctx->h_fctl};
    ctx->tmp_self.release();
    ctx->tmp_cand.release();
    ctx->ref_plan.release();
    for (DevBuf* b : bufs) b->release();
    ctx->pin_a.release();
    ctx->pin_b.release();
    ctx->pin_c.release();
    ctx->h_misc.release();
    ctx->h_groups.release();
    peer_release(ctx);
    ctx->pin_frames[0].release();
    ctx->pin_frames[1].release();
    free_coef_cache(ctx);
    for (int k = 0; k < 2; ++k) {
        if (ctx->ev_copy[k]) cudaEventDestroy(ctx->ev_copy[k]);
        if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
        if (ctx->ev_chunk[2 * k]) cudaEventDestroy(ctx->ev_chunk[2 * k]);
        if (ctx->ev_chunk[2 * k + 1]) cudaEventDestroy(ctx->ev_chunk[2 * k + 1]);
    }
    for (int k = 0; k < 4; ++k) {
        if (ctx->kt0[k]) cudaEventDestroy(ctx->kt0[k]);
        if (ctx->kt1[k]) cudaEventDestroy(ctx->kt1[k]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->lb_stream) cudaStreamDestroy(ctx->lb_stream);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    delete ctx;
}

const char* vdf_last_error(const vdf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vdf_ctx_set_shard(vdf_ctx* ctx, uint32_t rank, uint32_t world) {
    if (!ctx) return VDF_ERR_INVALID;
    if (world == 0 || rank >= world) {
        ctx->err = "rank must be < world";
        return VDF_ERR_INVALID;
    }
    ctx->rank = rank;
    ctx->world = world;
    return VDF_OK;
}

int vdf_ctx_set_option(vdf_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return VDF_ERR_INVALID;
    const std::string k(key);
    for (int r = 1; r < ctx->sub_count; ++r)  // every device of a multi-GPU context runs with the same knobs
        if (k != "exchange") VDF_TRY(vdf_ctx_set_option(ctx->sub[r], key, value));
    // the grouping indexes edges with 32 bits (group.cu)
    if (k == "max_edges" && value > 0) ctx->max_edges = std::min<uint64_t>((uint64_t)value, 0xFFFFFFFEull);
    else if (k == "initial_edges" && value > 0) ctx->initial_edges = (uint64_t)value;
    else if (k == "search_variant" && ((value >= 0 && value <= 2) || value == 5 || value == 6)) ctx->search_variant = (int)value;
    else if (k == "tc_fold" && (value == -1 || value == 0)) ctx->tc_fold = (int)value;
    else if (k == "peer_timeout_ms" && value >= 0) ctx->peer_timeout_ms = (uint64_t)value;
    else if (k == "tc_chunk" && value >= 0 && value <= 65535) ctx->tc_chunk = (uint32_t)value;
    else if (k == "tc_unit_order" && (value == 0 || value == 1)) ctx->tc_unit_order = (uint32_t)value;
    else if (k == "tc_a_tmem" && (value == 0 || value == 1)) ctx->tc_a_tmem = (uint32_t)value;
    else if (k == "hash_chunks" && value >= 1 && value <= 4) ctx->hash_chunks = (uint32_t)value;
    else if (k == "hash_overlap" && (value == 0 || value == 1)) ctx->hash_overlap = (uint32_t)value;
    else if (k == "hash_fuse_dct" && (value == 0 || value == 1)) ctx->hash_fuse_dct = (uint32_t)value;
    else if (k == "hash_fused" && (value == 0 || value == 1)) ctx->hash_fused = (uint32_t)value;
    else if (k == "grouping" && (value == 0 || value == 1)) ctx->grouping = (int)value;
    else if (k == "exchange" && (value == 0 || value == 1)) ctx->exchange = (int)value;
    else if (k == "hash_variant" && value >= 0 && value <= 3) ctx->hash_variant = (int)value;
    else {
        ctx->err = "unknown option or bad value: " + k;
        return VDF_ERR_INVALID;
    }
    return VDF_OK;
}

// ---- edge exchange over peer memory (common.cuh: PeerExchange) ----------------------------------------------------
}  // extern "C"
void vdf::peer_release(vdf_ctx* ctx) {
    PeerExchange& px = ctx->peer;
    for (uint32_t r = 0; r < px.world; ++r)
        if (px.ipc && r != px.rank && px.mapped[r]) cudaIpcCloseMemHandle(px.mapped[r]);
    if (px.local) cudaFree(px.local);
    px = PeerExchange();
    ctx->exchange = 0;
    ctx->peer_dead = false;
}
extern "C" {

int vdf_peer_alloc(vdf_ctx* ctx, uint64_t capacity_keys, unsigned char handle_out[64]) {
    if (!ctx || !handle_out || capacity_keys == 0) return VDF_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    ctx->err.clear();
    VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    peer_release(ctx);
    PeerExchange& px = ctx->peer;
    px.cap = (capacity_keys + 31) / 32 * 32;
    const size_t bytes = 2 * (256 + px.cap * 8);
    VDF_ALLOC(ctx, cudaMalloc(&px.local, bytes));
    VDF_CUDA(ctx, cudaMemset(px.local, 0, bytes));
    cudaIpcMemHandle_t h;
    VDF_CUDA(ctx, cudaIpcGetMemHandle(&h, px.local));
    memcpy(handle_out, &h, 64);
    return VDF_OK;
}

int vdf_peer_open(vdf_ctx* ctx, uint32_t rank, uint32_t world, const unsigned char* handles) {
    if (!ctx || !handles) return VDF_ERR_INVALID;
    ctx->err.clear();
    PeerExchange& px = ctx->peer;
    if (!px.local || px.world) {
        ctx->err = "vdf_peer_open: call vdf_peer_alloc first (once per vdf_peer_open)";
        return VDF_ERR_INVALID;
    }
    if (world < 2 || world > (uint32_t)kMaxPeers || rank >= world) {
        ctx->err = "vdf_peer_open: 2 <= world <= 8 GPUs of one node, rank < world";
        return VDF_ERR_INVALID;
    }
    VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    for (uint32_t r = 0; r < world; ++r) {
        if (r == rank) {
            px.mapped[r] = px.local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&px.mapped[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            for (uint32_t q = 0; q < r; ++q)
                if (q != rank && px.mapped[q]) cudaIpcCloseMemHandle(px.mapped[q]), px.mapped[q] = nullptr;
            return VDF_ERR_CUDA;
        }
    }
    px.rank = rank, px.world = world, px.epoch = 0, px.ipc = true;
    return VDF_OK;
}

int vdf_peer_close(vdf_ctx* ctx) {
    if (!ctx) return VDF_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    peer_release(ctx);
    return VDF_OK;
}

void* vdf_ctx_stream(vdf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void vdf_ctx_counters(const vdf_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!ctx) return;
    if (kernel_launches) *kernel_launches = ctx->launches;
    if (h2d_bytes) *h2d_bytes = ctx->h2d;
    if (d2h_bytes) *d2h_bytes = ctx->d2h;
}

int vdf_ctx_kernel_time(vdf_ctx* ctx, int which, double* total_ms, uint64_t* launches, int reset) {
    if (!ctx || which < 0 || which > 3) return VDF_ERR_INVALID;
    kt_collect(ctx);
    if (total_ms) *total_ms = ctx->kt_ms[which];
    if (launches) *launches = ctx->kt_n[which];
    if (reset) ctx->kt_ms[which] = 0, ctx->kt_n[which] = 0;
    return VDF_OK;
}

void vdf_free_edges(vdf_edges* e) {
    if (!e) return;
    free(e->ij);
    e->ij = nullptr;
    e->n = 0;
}
void vdf_free_groups(vdf_groups* g) {
    if (!g) return;
    // one block (group_ptr, then member_idx) from the library's result pool, see group.cu
    if (g->group_ptr && g->member_idx && g->member_idx != g->group_ptr + g->n_groups + 1) free(g->member_idx);
    vdf::result_free(g->group_ptr);
    g->group_ptr = g->member_idx = nullptr;
    g->n_groups = 0;
}
void vdf_free_csr(vdf_csr* c) {
    if (!c) return;
    free(c->row_ptr);
    free(c->col_idx);
    c->row_ptr = c->col_idx = nullptr;
    c->n_rows = 0;
}

}  // extern "C"

namespace vdf {
void kt_collect(vdf_ctx* ctx) {
    for (int k = 0; k < 4; ++k) {
        if (!ctx->kt_pending[k]) continue;
        if (cudaEventSynchronize(ctx->kt1[k]) == cudaSuccess) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->kt0[k], ctx->kt1[k]) == cudaSuccess) ctx->kt_ms[k] += ms, ctx->kt_n[k]++;
        }
        cudaGetLastError();
        ctx->kt_pending[k] = false;
    }
}
void kt_begin(vdf_ctx* ctx, int which, cudaStream_t stream) {
    if (ctx->kt_pending[which]) kt_collect(ctx);
    cudaEventRecord(ctx->kt0[which], stream ? stream : ctx->stream);
}
void kt_end(vdf_ctx* ctx, int which, cudaStream_t stream) {
    cudaEventRecord(ctx->kt1[which], stream ? stream : ctx->stream);
    ctx->kt_pending[which] = true;
}
}  // namespace vdf

// ------------------------------------------------------------------------------------------------ helpers
static int enter(vdf_ctx* ctx) {
    if (!ctx) return VDF_ERR_INVALID;
    ctx->err.clear();
    if (cudaSetDevice(ctx->device) != cudaSuccess) {
        ctx->err = "cudaSetDevice failed";
        return VDF_ERR_CUDA;
    }
    cudaGetLastError();  // an error a previous call already reported must not be picked up by this one's launch checks
    return VDF_OK;
}

static int upload(vdf_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
    VDF_ALLOC(ctx, buf.ensure(bytes ? bytes : 16));
    if (bytes) VDF_CUDA(ctx, cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += bytes;
    return VDF_OK;
}

// run a device search that writes sorted keys, growing the key buffer if the match count exceeds it
template <typename F>
static int with_growing_keys(vdf_ctx* ctx, F&& run, uint64_t* n_out) {
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
            ctx->err = "edge buffer overflow: " + std::to_string(cnt) + " matches exceed max_edges " +
                       std::to_string(ctx->max_edges);
            return VDF_ERR_EDGE_OVERFLOW;
        }
        cap = cnt + cnt / 16 + 1024;  // the count is exact for the same inputs: one retry suffices
    }
    return VDF_ERR_EDGE_OVERFLOW;
}

extern "C" {

// ------------------------------------------------------------------------------------------------ search path
int vdf_search_self_device(vdf_ctx* ctx, const uint64_t* d_hash_sorted, const uint32_t* d_dur_sorted, uint64_t n,
                           uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    VDF_TRY(enter(ctx));
    if (!n_out || (n && (!d_hash_sorted || !d_dur_sorted)) || (capacity && !d_keys_out)) return VDF_ERR_INVALID;
    int rc = search_self_device(ctx, d_hash_sorted, d_dur_sorted, n, tol_int, d_keys_out, capacity, n_out);
    cudaStreamSynchronize(ctx->stream);  // d_keys_out is read by the caller's own stream next
    return rc;
}

int vdf_search_refs_device(vdf_ctx* ctx, const uint64_t* d_cand_sorted, const uint32_t* d_cand_dur_sorted,
                           uint64_t n_cand, uint64_t cand_index_base, const uint64_t* d_refs,
                           const uint32_t* d_ref_dur, uint64_t n_ref, uint32_t tol_int, uint64_t* d_keys_out,
                           uint64_t capacity, uint64_t* n_out) {
    VDF_TRY(enter(ctx));
    if (!n_out || (capacity && !d_keys_out)) return VDF_ERR_INVALID;
    int rc = search_refs_device(ctx, d_cand_sorted, d_cand_dur_sorted, n_cand, cand_index_base, d_refs, d_ref_dur, n_ref,
                                tol_int, d_keys_out, capacity, n_out);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

int vdf_group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, const uint32_t* d_remap,
                            vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out) return VDF_ERR_INVALID;
    return group_device(ctx, n, d_keys_sorted, n_edges, d_remap, out);
}

// ---- prepared tables: `Search::from(hashes)` once, `search_self(tolerance)` / `search_with_references` many times ----
int vdf_table_create_device(vdf_ctx* ctx, const uint64_t* d_hash_sorted, const uint32_t* d_dur_sorted, uint64_t n, vdf_table** out) {
    VDF_TRY(enter(ctx));
    if (!out || (n && (!d_hash_sorted || !d_dur_sorted))) return VDF_ERR_INVALID;
    *out = nullptr;
    if (n >= 0xFFFFFF00ull) {
        ctx->err = "n must be < 2^32";
        return VDF_ERR_INVALID;
    }
    vdf_table* t = new (std::nothrow) vdf_table();
    if (!t) return VDF_ERR_ALLOC;
    t->ctx = ctx;
    int rc = table_prepare(ctx, t->t, d_hash_sorted, nullptr, d_dur_sorted, n, true, true);
    if (rc == VDF_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) ctx->err = "vdf_table_create: packing failed", rc = VDF_ERR_CUDA;
    if (rc != VDF_OK) {
        t->t.release();
        delete t;
        return rc;
    }
    *out = t;
    return VDF_OK;
}

int vdf_table_create(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n, vdf_table** out) {
    VDF_TRY(enter(ctx));
    if (!out || (n && (!hash_sorted || !dur_sorted))) return VDF_ERR_INVALID;
    *out = nullptr;
    vdf_table* t = new (std::nothrow) vdf_table();
    if (!t) return VDF_ERR_ALLOC;
    t->ctx = ctx;
    int rc = upload(ctx, t->own_hash, hash_sorted, n * 128);
    if (rc == VDF_OK) rc = upload(ctx, t->own_dur, dur_sorted, n * 4);
    if (rc == VDF_OK) rc = table_prepare(ctx, t->t, t->own_hash.as<uint64_t>(), nullptr, t->own_dur.as<uint32_t>(), n, true, true);
    if (rc == VDF_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) ctx->err = "vdf_table_create: upload failed", rc = VDF_ERR_CUDA;
    if (rc != VDF_OK) {
        t->t.release(), t->own_hash.release(), t->own_dur.release();
        delete t;
        return rc;
    }
    *out = t;
    return VDF_OK;
}

void vdf_table_destroy(vdf_table* t) {
    if (!t) return;
    if (t->ctx) {
        cudaSetDevice(t->ctx->device);
        cudaStreamSynchronize(t->ctx->stream);
    }
    t->t.release();
    t->own_hash.release();
    t->own_dur.release();
    delete t;
}

uint64_t vdf_table_len(const vdf_table* t) { return t ? t->t.n : 0; }

int vdf_table_search_self_device(vdf_table* t, uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    if (!t) return VDF_ERR_INVALID;
    VDF_TRY(enter(t->ctx));
    if (!n_out || (capacity && !d_keys_out)) return VDF_ERR_INVALID;
    int rc = table_search_self(t->ctx, t->t, tol_int, d_keys_out, capacity, n_out);
    cudaStreamSynchronize(t->ctx->stream);  // d_keys_out is read by the caller's own stream next
    return rc;
}

int vdf_table_search_self_groups(vdf_table* t, uint32_t tol_int, vdf_groups* out) {
    if (!t) return VDF_ERR_INVALID;
    vdf_ctx* ctx = t->ctx;
    VDF_TRY(enter(ctx));
    if (!out) return VDF_ERR_INVALID;
    uint64_t ne = 0;
    VDF_TRY(with_growing_keys(
        ctx, [&](uint64_t* keys, uint64_t cap, uint64_t* cnt) { return table_search_self(ctx, t->t, tol_int, keys, cap, cnt); }, &ne));
    return group_device(ctx, t->t.n, ctx->keys_a.as<uint64_t>(), ne, nullptr, out);
}

int vdf_table_search_refs_device(vdf_table* cand, uint64_t cand_index_base, const uint64_t* d_refs, const uint32_t* d_ref_dur, uint64_t n_ref,
                                 uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out) {
    if (!cand) return VDF_ERR_INVALID;
    VDF_TRY(enter(cand->ctx));
    if (!n_out || (capacity && !d_keys_out) || (n_ref && (!d_refs || !d_ref_dur))) return VDF_ERR_INVALID;
    int rc = table_search_refs(cand->ctx, cand->t, cand_index_base, d_refs, d_ref_dur, n_ref, tol_int, d_keys_out, capacity, n_out);
    cudaStreamSynchronize(cand->ctx->stream);
    return rc;
}

static int self_keys_from_host(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n,
                               uint32_t tol_int, uint64_t* n_keys) {
    VDF_TRY(upload(ctx, ctx->in_hash, hash_sorted, n * 128));
    VDF_TRY(upload(ctx, ctx->in_dur, dur_sorted, n * 4));
    return with_growing_keys(
        ctx,
        [&](uint64_t* keys, uint64_t cap, uint64_t* cnt) {
            return search_self_device(ctx, ctx->in_hash.as<uint64_t>(), ctx->in_dur.as<uint32_t>(), n, tol_int, keys, cap, cnt);
        },
        n_keys);
}

int vdf_search_self(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n, uint32_t tol_int,
                    vdf_edges* out) {
    VDF_TRY(enter(ctx));
    if (!out || (n && (!hash_sorted || !dur_sorted))) return VDF_ERR_INVALID;
    out->n = 0;
    out->ij = nullptr;
    uint64_t ne = 0;
    if (n) VDF_TRY(self_keys_from_host(ctx, hash_sorted, dur_sorted, n, tol_int, &ne));
    out->ij = (uint64_t*)malloc((ne ? ne : 1) * 16);
    if (!out->ij) return VDF_ERR_ALLOC;
    if (ne) {
        uint64_t* keys = (uint64_t*)malloc(ne * 8);
        if (!keys) return VDF_ERR_ALLOC;
        cudaError_t e = cudaMemcpyAsync(keys, ctx->keys_a.p, ne * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            free(keys);
            ctx->err = cudaGetErrorString(e);
            return VDF_ERR_CUDA;
        }
        ctx->d2h += ne * 8;
        for (uint64_t k = 0; k < ne; ++k) {
            out->ij[2 * k] = keys[k] >> 32;
            out->ij[2 * k + 1] = keys[k] & 0xFFFFFFFFull;
        }
        free(keys);
    }
    out->n = ne;
    return VDF_OK;
}

int vdf_group_greedy(vdf_ctx* ctx, uint64_t n, const vdf_edges* edges, vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out || !edges || (edges->n && !edges->ij)) return VDF_ERR_INVALID;
    const uint64_t ne = edges->n;
    std::vector<uint64_t> keys(ne);
    for (uint64_t k = 0; k < ne; ++k) {
        const uint64_t i = edges->ij[2 * k], j = edges->ij[2 * k + 1];
        if (i >= j || j >= n) {
            ctx->err = "edge list must hold i < j < n";
            return VDF_ERR_INVALID;
        }
        keys[k] = (i << 32) | j;
    }
    if (!std::is_sorted(keys.begin(), keys.end())) std::sort(keys.begin(), keys.end());
    VDF_TRY(upload(ctx, ctx->keys_b, keys.data(), ne * 8));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return group_greedy_device(ctx, n, ctx->keys_b.as<uint64_t>(), ne, nullptr, out);
}

int vdf_search_self_groups(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n,
                           uint32_t tol_int, vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out || (n && (!hash_sorted || !dur_sorted))) return VDF_ERR_INVALID;
    uint64_t ne = 0;
    if (n) VDF_TRY(self_keys_from_host(ctx, hash_sorted, dur_sorted, n, tol_int, &ne));
    return group_device(ctx, n, ctx->keys_a.as<uint64_t>(), ne, nullptr, out);
}

static int edges_to_device(vdf_ctx* ctx, uint64_t n, const vdf_edges* edges, uint64_t* ne_out) {
    const uint64_t ne = edges->n;
    std::vector<uint64_t> keys(ne);
    for (uint64_t k = 0; k < ne; ++k) {
        const uint64_t i = edges->ij[2 * k], j = edges->ij[2 * k + 1];
        if (i >= j || j >= n) {
            ctx->err = "edge list must hold i < j < n";
            return VDF_ERR_INVALID;
        }
        keys[k] = (i << 32) | j;
    }
    if (!std::is_sorted(keys.begin(), keys.end())) std::sort(keys.begin(), keys.end());
    VDF_TRY(upload(ctx, ctx->keys_b, keys.data(), ne * 8));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *ne_out = ne;
    return VDF_OK;
}

int vdf_group_components(vdf_ctx* ctx, uint64_t n, const vdf_edges* edges, vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out || !edges || (edges->n && !edges->ij)) return VDF_ERR_INVALID;
    uint64_t ne = 0;
    VDF_TRY(edges_to_device(ctx, n, edges, &ne));
    return group_components_device(ctx, n, ctx->keys_b.as<uint64_t>(), ne, nullptr, out);
}

int vdf_group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t n_edges, vdf_groups* out) {
    VDF_TRY(enter(ctx));
    if (!out) return VDF_ERR_INVALID;
    return group_components_device(ctx, n, d_keys, n_edges, nullptr, out);
}

int vdf_search_refs(vdf_ctx* ctx, const uint64_t* cand_sorted, const uint32_t* cand_dur_sorted, uint64_t n_cand,
                    const uint64_t* refs, const uint32_t* ref_dur, uint64_t n_ref, uint32_t tol_int, vdf_csr* out) {
    VDF_TRY(enter(ctx));
    if (!out || (n_cand && (!cand_sorted || !cand_dur_sorted)) || (n_ref && (!refs || !ref_dur))) return VDF_ERR_INVALID;
    out->n_rows = n_ref;
    out->row_ptr = (uint64_t*)calloc(n_ref + 1, 8);
    out->col_idx = nullptr;
    if (!out->row_ptr) return VDF_ERR_ALLOC;
    uint64_t ne = 0;
    if (n_cand && n_ref) {
        VDF_TRY(upload(ctx, ctx->in_hash, cand_sorted, n_cand * 128));
        VDF_TRY(upload(ctx, ctx->in_dur, cand_dur_sorted, n_cand * 4));
        VDF_TRY(upload(ctx, ctx->in_hash2, refs, n_ref * 128));
        VDF_TRY(upload(ctx, ctx->in_dur2, ref_dur, n_ref * 4));
        VDF_TRY(with_growing_keys(
            ctx,
            [&](uint64_t* keys, uint64_t cap, uint64_t* cnt) {
                return search_refs_device(ctx, ctx->in_hash.as<uint64_t>(), ctx->in_dur.as<uint32_t>(), n_cand, 0,
                                          ctx->in_hash2.as<uint64_t>(), ctx->in_dur2.as<uint32_t>(), n_ref, tol_int, keys,
                                          cap, cnt);
            },
            &ne));
    }
    out->col_idx = (uint64_t*)malloc((ne ? ne : 1) * 8);
    if (!out->col_idx) return VDF_ERR_ALLOC;
    if (ne) {
        VDF_CUDA(ctx, cudaMemcpyAsync(out->col_idx, ctx->keys_a.p, ne * 8, cudaMemcpyDeviceToHost, ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2h += ne * 8;
        // keys are sorted by (ref, cand): count per ref, then strip the ref half
        for (uint64_t k = 0; k < ne; ++k) {
            out->row_ptr[(out->col_idx[k] >> 32) + 1]++;
            out->col_idx[k] &= 0xFFFFFFFFull;
        }
        for (uint64_t r = 0; r < n_ref; ++r) out->row_ptr[r + 1] += out->row_ptr[r];
    }
    return VDF_OK;
}

int vdf_self_window_pairs(vdf_ctx* ctx, const uint32_t* dur_sorted, uint64_t n, uint64_t* pairs_out) {
    VDF_TRY(enter(ctx));
    if (!pairs_out || (n && !dur_sorted)) return VDF_ERR_INVALID;
    VDF_TRY(upload(ctx, ctx->in_dur, dur_sorted, n * 4));
    return self_window_pairs(ctx, ctx->in_dur.as<uint32_t>(), n, pairs_out);
}

// ------------------------------------------------------------------------------------------------ hashing path
int vdf_hash_stacks_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                           uint64_t* d_out_hash, int32_t* out_status, uint32_t* out_crop) {
    VDF_TRY(enter(ctx));
    if (n && (!d_frames || !desc || !d_out_hash)) return VDF_ERR_INVALID;
    return hash_stacks_device(ctx, d_frames, desc, n, cropdetect, d_out_hash, nullptr, out_status, out_crop);
}

int vdf_hash_stacks_small_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n,
                                 int cropdetect, uint8_t* d_out_small, uint32_t* out_crop) {
    VDF_TRY(enter(ctx));
    if (n && (!d_frames || !desc || !d_out_small)) return VDF_ERR_INVALID;
    return hash_stacks_device(ctx, d_frames, desc, n, cropdetect, nullptr, d_out_small, nullptr, out_crop);
}

int vdf_hash_from_small(vdf_ctx* ctx, const uint8_t* small, uint32_t n, uint64_t* out_hash) {
    VDF_TRY(enter(ctx));
    if (n && (!small || !out_hash)) return VDF_ERR_INVALID;
    if (n == 0) return VDF_OK;
    VDF_TRY(upload(ctx, ctx->h_small, small, (size_t)n * 4096));
    VDF_ALLOC(ctx, ctx->h_hash.ensure((size_t)n * 128));
    VDF_TRY(hash_from_small_device(ctx, ctx->h_small.as<uint8_t>(), n, ctx->h_hash.as<uint64_t>()));
    VDF_CUDA(ctx, cudaMemcpyAsync(out_hash, ctx->h_hash.p, (size_t)n * 128, cudaMemcpyDeviceToHost, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d2h += (size_t)n * 128;
    return VDF_OK;
}

// Host frames: stacks are packed (pitch = width) into one of two HBM staging buffers by async copies from
// pinned memory on a copy stream, so that batch b+1 uploads while batch b is being hashed.
int vdf_hash_stacks(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                    uint64_t* out_hash, int32_t* out_status, uint32_t* out_crop) {
    VDF_TRY(enter(ctx));
    if (n && (!frames || !desc || !out_hash)) return VDF_ERR_INVALID;
    if (n == 0) return VDF_OK;
    if (ctx->sub_count > 1) return mgpu_hash_stacks(ctx, frames, desc, n, cropdetect, out_hash, out_status, out_crop);
    return hash_stacks_host(ctx, frames, desc, n, cropdetect, out_hash, out_status, out_crop);
}

}  // extern "C"

int vdf::hash_stacks_host(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect, uint64_t* out_hash,
                          int32_t* out_status, uint32_t* out_crop) {
    VDF_TRY(enter(ctx));
    const size_t kBatchBytes = (size_t)512 << 20;
    cudaPointerAttributes attr;
    bool src_pinned = false;
    if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess) src_pinned = (attr.type == cudaMemoryTypeHost);
    cudaGetLastError();

    auto stack_bytes = [&](const vdf_stack_desc& d) -> size_t {
        if ((d.flags & VDF_STACK_FLAG_MIXED_SIZES) || d.n_frames < VDF_DCT_SIZE) return 0;
        return (size_t)VDF_DCT_SIZE * d.width * d.height;
    };
    struct Batch {
        uint32_t first, count;
        size_t bytes;
    };
    std::vector<Batch> batches;
    for (uint32_t s = 0; s < n;) {
        Batch b{s, 0, 0};
        while (s < n && (b.count == 0 || b.bytes + stack_bytes(desc[s]) <= kBatchBytes)) {
            b.bytes += stack_bytes(desc[s]);
            ++b.count;
            ++s;
        }
        batches.push_back(b);
    }
    std::vector<std::vector<vdf_stack_desc>> packed(2);
    auto stage = [&](size_t bi) -> int {  // issue the upload of batch bi into staging buffer bi & 1
        const Batch& b = batches[bi];
        const int buf = (int)(bi & 1);
        VDF_ALLOC(ctx, ctx->h_frames[buf].ensure(b.bytes ? b.bytes : 16));
        if (!src_pinned) VDF_ALLOC(ctx, ctx->pin_frames[buf].ensure(b.bytes ? b.bytes : 16));
        VDF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[buf], 0));  // buffer free again?
        if (!src_pinned) VDF_CUDA(ctx, cudaEventSynchronize(ctx->ev_done[buf]));
        packed[buf].assign(desc + b.first, desc + b.first + b.count);
        size_t off = 0;
        uint8_t* dev = ctx->h_frames[buf].as<uint8_t>();
        for (uint32_t k = 0; k < b.count; ++k) {
            vdf_stack_desc& d = packed[buf][k];
            const size_t sb = stack_bytes(d);
            if (sb) {
                const uint8_t* src = frames + d.offset;
                const size_t fb = (size_t)d.width * d.height;
                const bool contiguous = d.pitch == d.width && d.frame_stride == fb;
                if (contiguous && src_pinned) {  // one 16-frame copy instead of 16 strided ones
                    VDF_CUDA(ctx, cudaMemcpyAsync(dev + off, src, sb, cudaMemcpyHostToDevice, ctx->copy_stream));
                } else if (contiguous) {
                    std::memcpy(ctx->pin_frames[buf].as<uint8_t>() + off, src, sb);
                } else
                for (uint32_t f = 0; f < VDF_DCT_SIZE; ++f) {
                    const uint8_t* fs = src + (size_t)f * d.frame_stride;
                    if (src_pinned) {
                        VDF_CUDA(ctx, cudaMemcpy2DAsync(dev + off + f * fb, d.width, fs, d.pitch, d.width, d.height,
                                                        cudaMemcpyHostToDevice, ctx->copy_stream));
                    } else {
                        uint8_t* pin = ctx->pin_frames[buf].as<uint8_t>() + off + f * fb;
                        if (d.pitch == d.width) std::memcpy(pin, fs, fb);
                        else
                            for (uint32_t y = 0; y < d.height; ++y)
                                std::memcpy(pin + (size_t)y * d.width, fs + (size_t)y * d.pitch, d.width);
                    }
                }
                d.offset = off;
                d.frame_stride = fb;
                d.pitch = d.width;
                off += sb;
            }
        }
        if (!src_pinned && b.bytes)
            VDF_CUDA(ctx, cudaMemcpyAsync(dev, ctx->pin_frames[buf].p, b.bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        ctx->h2d += b.bytes;
        VDF_CUDA(ctx, cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
        return VDF_OK;
    };
    VDF_TRY(stage(0));
    for (size_t bi = 0; bi < batches.size(); ++bi) {
        const Batch& b = batches[bi];
        const int buf = (int)(bi & 1);
        if (bi + 1 < batches.size()) VDF_TRY(stage(bi + 1));
        VDF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        VDF_ALLOC(ctx, ctx->h_hash.ensure((size_t)b.count * 128));
        VDF_TRY(hash_stacks_device(ctx, ctx->h_frames[buf].as<uint8_t>(), packed[buf].data(), b.count, cropdetect,
                                   ctx->h_hash.as<uint64_t>(), nullptr, out_status ? out_status + b.first : nullptr,
                                   out_crop ? out_crop + (size_t)b.first * 4 : nullptr));
        VDF_CUDA(ctx, cudaMemcpyAsync(out_hash + (size_t)b.first * 16, ctx->h_hash.p, (size_t)b.count * 128,
                                      cudaMemcpyDeviceToHost, ctx->stream));
        VDF_CUDA(ctx, cudaEventRecord(ctx->ev_done[buf], ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2h += (size_t)b.count * 128;
    }
    return VDF_OK;
}
