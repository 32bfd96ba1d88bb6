// mgpu.cu -- several GPUs of one node behind ONE context ("a ctx owns its devices and comms").
//
// vdf_ctx_create_multi(dev_ids, n_dev) returns a context that vdf_search, vdf_search_with_references and vdf_hash_stacks take
// like any other: a caller of the crate's `search()` (vid_dup_finder_app/src/app/app_fns.rs:478-482) uses 8 GPUs with one call.
// One process, one host thread per device while a call runs; peer access is enabled directly (no IPC):
//   * search: the caller's table is staged and sorted ONCE on device 0 (host.cu: stage_and_sort), copied GPU -> GPU over
//     NVLink, every device packs it and evaluates its block-cyclic share of the (row pair, chunk) units; matches travel
//     through the fused peer exchange of the pair kernel (search_tc.cu: tc6_emit, search.cu: peer_barrier_kernel), so
//     every device ends up with the whole edge list; the greedy grouping runs on device 0.
//   * search_with_references: contiguous slices of the sorted candidate table per device, references replicated.
//   * hashing: contiguous shards of stacks, no exchange.
// The one-process-per-GPU plane (torch.distributed, vid_dup_finder_lib_b200/dist.py) stays for callers that already are SPMD.
#include <algorithm>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace vdf {
namespace {

template <typename F>
int on_all(vdf_ctx* ctx, F&& body) {  // body(rank, sub-context) -> rc, one host thread per device
    const int W = ctx->sub_count;
    std::vector<int> rcs(W, VDF_OK);
    std::vector<std::thread> th;
    th.reserve(W);
    for (int r = 0; r < W; ++r)
        th.emplace_back([&, r] {
            cudaSetDevice(ctx->sub[r]->device);
            rcs[r] = body(r, ctx->sub[r]);
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < W; ++r)
        if (rcs[r] != VDF_OK && rcs[r] != VDF_ERR_EDGE_OVERFLOW) {
            if (r) ctx->err = "device " + std::to_string(ctx->sub[r]->device) + ": " + ctx->sub[r]->err;
            return rcs[r];
        }
    for (int r = 0; r < W; ++r)
        if (rcs[r] == VDF_ERR_EDGE_OVERFLOW) {
            if (r) ctx->err = ctx->sub[r]->err;
            return rcs[r];
        }
    return VDF_OK;
}

// (re)allocate the exchange buffer of every device for `capacity` keys in total and cross-map them (same process: a peer's
// cudaMalloc pointer is valid on every device that enabled access to it)
int exchange_setup(vdf_ctx* ctx, uint64_t capacity) {
    const int W = ctx->sub_count;
    for (int r = 0; r < W; ++r) {
        vdf_ctx* s = ctx->sub[r];
        VDF_CUDA(ctx, cudaSetDevice(s->device));
        VDF_CUDA(ctx, cudaStreamSynchronize(s->stream));
        peer_release(s);
        PeerExchange& px = s->peer;
        px.cap = (capacity + 31) / 32 * 32;
        const size_t bytes = 2 * (256 + px.cap * 8);
        VDF_ALLOC(ctx, cudaMalloc(&px.local, bytes));
        VDF_CUDA(ctx, cudaMemset(px.local, 0, bytes));
    }
    for (int r = 0; r < W; ++r) {
        PeerExchange& px = ctx->sub[r]->peer;
        for (int q = 0; q < W; ++q) px.mapped[q] = ctx->sub[q]->peer.local;
        px.rank = (uint32_t)r, px.world = (uint32_t)W, px.epoch = 0, px.ipc = false;
    }
    VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    return VDF_OK;
}

// device 0 -> every other device, on the destination's stream, after device 0's stream reached `ready`
int fan_out(vdf_ctx* ctx, cudaEvent_t ready, const void* src, size_t bytes, DevBuf vdf_ctx::*dst) {
    for (int r = 1; r < ctx->sub_count; ++r) {
        vdf_ctx* s = ctx->sub[r];
        VDF_CUDA(ctx, cudaSetDevice(s->device));
        VDF_ALLOC(ctx, (s->*dst).ensure(bytes ? bytes : 16));
        VDF_CUDA(ctx, cudaStreamWaitEvent(s->stream, ready, 0));
        if (bytes) VDF_CUDA(ctx, cudaMemcpyPeerAsync((s->*dst).p, s->device, src, ctx->device, bytes, s->stream));
    }
    VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    return VDF_OK;
}

// run `search` on every device with the exchange on, growing the exchange buffers together until the matches fit
template <typename F>
int exchange_search(vdf_ctx* ctx, F&& search, uint64_t* n_keys) {
    const int W = ctx->sub_count;
    uint64_t cap = std::max<uint64_t>(ctx->initial_edges, 1024);
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (!ctx->peer.world || ctx->peer.cap < cap || ctx->peer_dead) VDF_TRY(exchange_setup(ctx, cap));
        cap = ctx->peer.cap;
        std::vector<uint64_t> cnt(W, 0);
        const int rc = on_all(ctx, [&](int r, vdf_ctx* s) {
            s->err.clear();
            if (s->keys_a.ensure(cap * 8) != cudaSuccess) {
                s->err = "key buffer allocation failed";
                cudaGetLastError();
                return VDF_ERR_ALLOC;
            }
            s->exchange = 1;
            const int rc1 = search(r, s, s->keys_a.as<uint64_t>(), cap, &cnt[r]);
            s->exchange = 0;
            return rc1;
        });
        if (rc == VDF_OK) {
            *n_keys = cnt[0];
            return VDF_OK;
        }
        if (rc != VDF_ERR_EDGE_OVERFLOW) return rc;
        uint64_t need = 0;
        for (int r = 0; r < W; ++r) need = std::max(need, cnt[r]);
        if (need > ctx->max_edges) {
            ctx->err = "edge buffer overflow: " + std::to_string(need) + " matches exceed max_edges " + std::to_string(ctx->max_edges);
            return VDF_ERR_EDGE_OVERFLOW;
        }
        cap = need + need / 16 + 1024;
    }
    return VDF_ERR_EDGE_OVERFLOW;
}

}  // namespace

int mgpu_search(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n,
                double tolerance, vdf_groups* out) {
    const int W = ctx->sub_count;
    if (n == 0) return group_device(ctx, 0, nullptr, 0, nullptr, out);
    double t_keys = 0, t_stage = 0;
    VDF_TRY(stage_and_sort(ctx, hashes, durations, path_blob, path_off, n, ctx->in_hash, ctx->pin_a, &t_keys, &t_stage));
    VDF_CUDA(ctx, cudaEventRecord(ctx->ev_copy[0], ctx->stream));
    VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_hash.p, n * 128, &vdf_ctx::in_hash));
    VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->sk_order.p, n * 4, &vdf_ctx::sk_order));
    VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_dur.p, n * 4, &vdf_ctx::in_dur));
    const uint32_t tol_int = tolerance_to_int(tolerance);
    uint64_t ne = 0;
    VDF_TRY(exchange_search(
        ctx,
        [&](int r, vdf_ctx* s, uint64_t* keys, uint64_t cap, uint64_t* cnt) -> int {
            s->rank = (uint32_t)r, s->world = (uint32_t)W;
            // Search::from on every device: gathered into sorted order while it is packed
            VDF_TRY(table_prepare(s, s->tmp_self, s->in_hash.as<uint64_t>(), s->sk_order.as<uint32_t>(), s->in_dur.as<uint32_t>(), n, true, false));
            const int rc = table_search_self(s, s->tmp_self, tol_int, keys, cap, cnt);
            s->rank = 0, s->world = 1;
            return rc;
        },
        &ne));
    VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    VDF_TRY(group_device(ctx, n, ctx->keys_a.as<uint64_t>(), ne, ctx->sk_order.as<uint32_t>(), out));
    ctx->phase_ms[0] = t_keys, ctx->phase_ms[1] = t_stage, ctx->phase_ms[2] = 0, ctx->phase_ms[3] = 0;
    return VDF_OK;
}

int mgpu_search_refs(vdf_ctx* ctx, const uint64_t* ref_hashes, const uint32_t* ref_durations, uint64_t n_ref, const uint64_t* cand_hashes,
                     const uint32_t* cand_durations, const char* cand_path_blob, const uint64_t* cand_path_off, uint64_t n_cand,
                     double tolerance, vdf_csr* out) {
    const int W = ctx->sub_count;
    out->n_rows = n_ref;
    out->row_ptr = (uint64_t*)calloc(n_ref + 1, 8);
    out->col_idx = nullptr;
    if (!out->row_ptr) return VDF_ERR_ALLOC;
    uint64_t ne = 0;
    if (n_cand && n_ref) {
        double t_keys = 0, t_stage = 0;
        VDF_TRY(stage_and_sort(ctx, cand_hashes, cand_durations, cand_path_blob, cand_path_off, n_cand, ctx->in_hash, ctx->pin_a, &t_keys, &t_stage));
        VDF_ALLOC(ctx, ctx->in_hash2.ensure(n_ref * 128));
        VDF_ALLOC(ctx, ctx->in_dur2.ensure(n_ref * 4));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_hash2.p, ref_hashes, n_ref * 128, cudaMemcpyHostToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->in_dur2.p, ref_durations, n_ref * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += n_ref * 132;
        VDF_CUDA(ctx, cudaEventRecord(ctx->ev_copy[0], ctx->stream));
        VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_hash.p, n_cand * 128, &vdf_ctx::in_hash));
        VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->sk_order.p, n_cand * 4, &vdf_ctx::sk_order));
        VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_dur.p, n_cand * 4, &vdf_ctx::in_dur));
        VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_hash2.p, n_ref * 128, &vdf_ctx::in_hash2));
        VDF_TRY(fan_out(ctx, ctx->ev_copy[0], ctx->in_dur2.p, n_ref * 4, &vdf_ctx::in_dur2));
        const uint32_t tol_int = tolerance_to_int(tolerance);
        VDF_TRY(exchange_search(
            ctx,
            [&](int r, vdf_ctx* s, uint64_t* keys, uint64_t cap, uint64_t* cnt) -> int {
                // contiguous slice [b, e) of the SORTED candidate table: per-reference lists concatenate in sorted order
                const uint64_t base = n_cand / W, rem = n_cand % W;
                const uint64_t b = r * base + std::min<uint64_t>(r, rem), e = b + base + ((uint64_t)r < rem ? 1 : 0);
                VDF_TRY(table_prepare(s, s->tmp_cand, s->in_hash.as<uint64_t>(), s->sk_order.as<uint32_t>() + b, s->in_dur.as<uint32_t>() + b, e - b,
                                      false, true));
                return table_search_refs(s, s->tmp_cand, b, s->in_hash2.as<uint64_t>(), s->in_dur2.as<uint32_t>(), n_ref, tol_int, keys, cap, cnt);
            },
            &ne));
        VDF_CUDA(ctx, cudaSetDevice(ctx->device));
    }
    return ref_keys_to_csr(ctx, ctx->keys_a.as<uint64_t>(), ne, ctx->sk_order.as<uint32_t>(), n_ref, out);
}

int mgpu_hash_stacks(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect, uint64_t* out_hash,
                     int32_t* out_status, uint32_t* out_crop) {
    const uint32_t W = (uint32_t)ctx->sub_count;
    return on_all(ctx, [&](int r, vdf_ctx* s) -> int {
        const uint32_t base = n / W, rem = n % W;
        const uint32_t b = r * base + std::min<uint32_t>(r, rem), cnt = base + ((uint32_t)r < rem ? 1 : 0);
        if (cnt == 0) return VDF_OK;
        return hash_stacks_host(s, frames, desc + b, cnt, cropdetect, out_hash + (size_t)b * 16, out_status ? out_status + b : nullptr,
                                out_crop ? out_crop + (size_t)b * 4 : nullptr);
    });
}

}  // namespace vdf

extern "C" {

int vdf_ctx_create_multi(const int* dev_ids, int n_dev, vdf_ctx** out) {
    if (!out || !dev_ids || n_dev < 1 || n_dev > vdf::kMaxPeers) return VDF_ERR_INVALID;
    *out = nullptr;
    for (int a = 0; a < n_dev; ++a)
        for (int b = a + 1; b < n_dev; ++b)
            if (dev_ids[a] == dev_ids[b]) return VDF_ERR_INVALID;
    vdf_ctx* ctx = nullptr;
    int rc = vdf_ctx_create(dev_ids[0], &ctx);
    if (rc != VDF_OK) return rc;
    ctx->sub[0] = ctx;
    ctx->sub_count = 1;
    for (int r = 1; r < n_dev; ++r) {
        rc = vdf_ctx_create(dev_ids[r], &ctx->sub[r]);
        if (rc != VDF_OK) {
            vdf_ctx_destroy(ctx);
            return rc;
        }
        ctx->sub_count = r + 1;
    }
    // every device reads and writes every other device's exchange buffer and receives the table from device 0
    for (int a = 0; a < n_dev && rc == VDF_OK; ++a) {
        cudaSetDevice(dev_ids[a]);
        for (int b = 0; b < n_dev; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, dev_ids[a], dev_ids[b]);
            const cudaError_t e = can ? cudaDeviceEnablePeerAccess(dev_ids[b], 0) : cudaErrorPeerAccessUnsupported;
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = VDF_ERR_NO_DEVICE;
            cudaGetLastError();
        }
    }
    cudaSetDevice(dev_ids[0]);
    if (rc != VDF_OK) {
        vdf_ctx_destroy(ctx);
        return rc;
    }
    if (n_dev == 1) ctx->sub_count = 0, ctx->sub[0] = nullptr;  // one device: an ordinary context
    *out = ctx;
    return VDF_OK;
}

int vdf_ctx_device_count(const vdf_ctx* ctx) { return ctx ? (ctx->sub_count > 1 ? ctx->sub_count : 1) : 0; }

}  // extern "C"
