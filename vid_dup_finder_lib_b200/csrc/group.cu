// group.cu -- the reference's greedy MatchGroup rule on the GPU.
//
// Replaces the consumption logic of Search::search_self (search_algorithm.rs:131-170): entries are visited in
// ascending order; an entry that nobody consumed becomes a *target* and consumes every still-unconsumed
// neighbour j > i.  Over the (i,j) edge list this is NOT connected components.  In closed form:
//   - v is a target  <=>  no neighbour u < v is a target            (lexicographically-first independent set)
//   - a non-target v joins its SMALLEST target neighbour u < v      (that target reaches v first)
// which parallelises as rounds over a worklist of undecided vertices: a vertex is decided once every smaller
// neighbour up to its first target (in ascending order) is decided.  Vertices without a smaller neighbour are
// targets from the start, so the number of rounds is the longest dependency chain (2-3 for duplicate clusters).
// Output order follows the reference: members ascending, the target last, groups by DESCENDING target
// (ret.reverse(), search_algorithm.rs:136,167).
#include <algorithm>
#include <cub/device/device_scan.cuh>

#include <cstring>

#include "common.cuh"

namespace vdf {

enum : uint8_t { kUndecided = 0, kTarget = 1, kMember = 2 };

__global__ void swap_halves_kernel(const uint64_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ out) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e < n) out[e] = (in[e] << 32) | (in[e] >> 32);
}

// rks = edges as (j << 32 | i), sorted: the in-edges (smaller neighbours, ascending) of each vertex j are contiguous.
// Every vertex with at least one in-edge starts undecided and enters the worklist (as the index of its segment).
__global__ void seed_worklist_kernel(const uint64_t* __restrict__ rks, uint64_t ne, uint8_t* __restrict__ state,
                                     uint32_t* __restrict__ wl, unsigned long long* __restrict__ count) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    uint32_t v = (uint32_t)(rks[e] >> 32);
    if (e == 0 || (uint32_t)(rks[e - 1] >> 32) != v) {
        state[v] = kUndecided;
        wl[atomicAdd(count, 1ull)] = (uint32_t)e;
    }
}

__global__ void greedy_round_kernel(const uint64_t* __restrict__ rks, uint64_t ne, volatile uint8_t* state,
                                    uint32_t* __restrict__ parent, const uint32_t* __restrict__ wl_in, uint64_t n_in,
                                    uint32_t* __restrict__ wl_out, unsigned long long* __restrict__ n_out) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= n_in) return;
    const uint32_t e0 = wl_in[k];
    const uint32_t v = (uint32_t)(rks[e0] >> 32);
    for (uint64_t e = e0; e < ne; ++e) {
        const uint64_t key = rks[e];
        if ((uint32_t)(key >> 32) != v) break;
        const uint32_t u = (uint32_t)key;
        const uint8_t s = state[u];
        if (s == kUndecided) {  // a smaller neighbour is still open: it may turn out to be the smallest target
            wl_out[atomicAdd(n_out, 1ull)] = e0;
            return;
        }
        if (s == kTarget) {  // everything before u is a member, so u is v's smallest target neighbour
            parent[v] = u;
            __threadfence();
            state[v] = kMember;
            return;
        }
    }
    state[v] = kTarget;
}

// one key per consumed vertex: (~target << 32 | vertex) so that an ascending sort lists groups by descending target
__global__ void member_keys_kernel(const uint64_t* __restrict__ rks, const uint8_t* __restrict__ state,
                                   const uint32_t* __restrict__ parent, const uint32_t* __restrict__ wl0, uint64_t nv,
                                   uint64_t* __restrict__ mk, unsigned long long* __restrict__ count) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= nv) return;
    const uint32_t v = (uint32_t)(rks[wl0[k]] >> 32);
    if (state[v] == kMember) mk[atomicAdd(count, 1ull)] = ((uint64_t)(~parent[v]) << 32) | v;
}

__global__ void group_flag_kernel(const uint64_t* __restrict__ mks, uint64_t nm, uint32_t* __restrict__ flag) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= nm) return;
    flag[k] = (k == 0 || (mks[k] >> 32) != (mks[k - 1] >> 32)) ? 1u : 0u;
}

// incl[k] = 1-based group number of member k.  Member k of group g lands at k + g; the target closes the group.
__global__ void group_csr_kernel(const uint64_t* __restrict__ mks, const uint32_t* __restrict__ incl, uint64_t nm,
                                 uint64_t* __restrict__ group_ptr, uint64_t* __restrict__ member_idx) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= nm) return;
    const uint64_t g = incl[k] - 1;
    member_idx[k + g] = (uint32_t)mks[k];
    const bool last = (k + 1 == nm) || (incl[k + 1] != incl[k]);
    if (last) {
        member_idx[k + g + 1] = (uint32_t)(~(uint32_t)(mks[k] >> 32));
        group_ptr[g + 1] = k + g + 2;
    }
    if (k == 0) group_ptr[0] = 0;
}

static int empty_groups(vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = (uint64_t*)calloc(1, sizeof(uint64_t));
    out->member_idx = (uint64_t*)calloc(1, sizeof(uint64_t));
    return (out->group_ptr && out->member_idx) ? VDF_OK : VDF_ERR_ALLOC;
}

// nm keys (~root << 32 | vertex) in g_mk -> sorted -> CSR (members ascending, the root last, groups by descending root)
// -> host arrays
static int finish_groups(vdf_ctx* ctx, unsigned long long nm, vdf_groups* out) {
    const int B = 256;
    auto blocks = [&](uint64_t items) { return (unsigned)((items + B - 1) / B); };
    cudaStream_t st = ctx->stream;
    if (nm == 0) return empty_groups(out);
    VDF_TRY(sort_keys(ctx, ctx->g_mk.as<uint64_t>(), ctx->g_mks.as<uint64_t>(), nm));

    VDF_ALLOC(ctx, ctx->g_flag.ensure(nm * 4));
    VDF_ALLOC(ctx, ctx->g_scan.ensure(nm * 4));
    VDF_ALLOC(ctx, ctx->g_gp.ensure((nm + 1) * 8));
    VDF_ALLOC(ctx, ctx->g_mem.ensure(2 * nm * 8));
    group_flag_kernel<<<blocks(nm), B, 0, st>>>(ctx->g_mks.as<uint64_t>(), nm, ctx->g_flag.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    size_t tmp = 0;
    VDF_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp, ctx->g_flag.as<uint32_t>(), ctx->g_scan.as<uint32_t>(),
                                                (size_t)nm, st));
    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
    VDF_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->sort_tmp.p, tmp, ctx->g_flag.as<uint32_t>(),
                                                ctx->g_scan.as<uint32_t>(), (size_t)nm, st));
    ctx->launches += 1;
    group_csr_kernel<<<blocks(nm), B, 0, st>>>(ctx->g_mks.as<uint64_t>(), ctx->g_scan.as<uint32_t>(), nm,
                                               ctx->g_gp.as<uint64_t>(), ctx->g_mem.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    uint32_t ng = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(&ng, ctx->g_scan.as<uint32_t>() + (nm - 1), 4, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));

    out->n_groups = ng;
    out->group_ptr = (uint64_t*)malloc((size_t)(ng + 1) * 8);
    out->member_idx = (uint64_t*)malloc((size_t)(nm + ng) * 8);
    if (!out->group_ptr || !out->member_idx) {
        ctx->err = "host allocation failed";
        return VDF_ERR_ALLOC;
    }
    // through pinned staging: a device -> pageable copy of these ~2 MB costs more than the rest of the grouping
    const size_t gp_bytes = (size_t)(ng + 1) * 8, mem_bytes = (size_t)(nm + ng) * 8;
    VDF_CUDA(ctx, ctx->h_groups.ensure(gp_bytes + mem_bytes));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->h_groups.p, ctx->g_gp.p, gp_bytes, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaMemcpyAsync(ctx->h_groups.as<uint8_t>() + gp_bytes, ctx->g_mem.p, mem_bytes, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    memcpy(out->group_ptr, ctx->h_groups.p, gp_bytes);
    memcpy(out->member_idx, ctx->h_groups.as<uint8_t>() + gp_bytes, mem_bytes);
    ctx->d2h += (size_t)(ng + 1) * 8 + (size_t)(nm + ng) * 8;
    return VDF_OK;
}

int group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t ne, vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = nullptr;
    out->member_idx = nullptr;
    if (ne == 0 || n == 0) return empty_groups(out);
    const int B = 256;
    auto blocks = [&](uint64_t items) { return (unsigned)((items + B - 1) / B); };
    cudaStream_t st = ctx->stream;

    VDF_ALLOC(ctx, ctx->g_rk.ensure(ne * 8));
    VDF_ALLOC(ctx, ctx->g_rks.ensure(ne * 8));
    VDF_ALLOC(ctx, ctx->g_state.ensure(n));
    VDF_ALLOC(ctx, ctx->g_parent.ensure(n * 4));
    VDF_ALLOC(ctx, ctx->g_wl0.ensure(ne * 4));
    VDF_ALLOC(ctx, ctx->g_wla.ensure(ne * 4));
    VDF_ALLOC(ctx, ctx->g_wlb.ensure(ne * 4));
    VDF_ALLOC(ctx, ctx->misc.ensure(64));
    unsigned long long* cnt = ctx->misc.as<unsigned long long>() + 4;  // [4] seeds, [5] ping, [6] pong, [7] members
    VDF_CUDA(ctx, cudaMemsetAsync(cnt, 0, 32, st));

    swap_halves_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, ctx->g_rk.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    VDF_TRY(sort_keys(ctx, ctx->g_rk.as<uint64_t>(), ctx->g_rks.as<uint64_t>(), ne));
    const uint64_t* rks = ctx->g_rks.as<uint64_t>();
    uint8_t* state = ctx->g_state.as<uint8_t>();
    VDF_CUDA(ctx, cudaMemsetAsync(state, kTarget, n, st));
    seed_worklist_kernel<<<blocks(ne), B, 0, st>>>(rks, ne, state, ctx->g_wl0.as<uint32_t>(), cnt + 0);
    VDF_LAUNCHED(ctx);
    unsigned long long nv = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(&nv, cnt + 0, 8, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));

    // rounds: ping-pong worklists until every vertex with a smaller neighbour is decided
    const uint32_t* wl_in = ctx->g_wl0.as<uint32_t>();
    uint32_t* bufs[2] = {ctx->g_wla.as<uint32_t>(), ctx->g_wlb.as<uint32_t>()};
    unsigned long long n_in = nv;
    for (int round = 0; n_in > 0; ++round) {
        unsigned long long* n_out = cnt + 1 + (round & 1);
        VDF_CUDA(ctx, cudaMemsetAsync(n_out, 0, 8, st));
        greedy_round_kernel<<<blocks(n_in), B, 0, st>>>(rks, ne, state, ctx->g_parent.as<uint32_t>(), wl_in, n_in,
                                                        bufs[round & 1], n_out);
        VDF_LAUNCHED(ctx);
        unsigned long long left = 0;
        VDF_CUDA(ctx, cudaMemcpyAsync(&left, n_out, 8, cudaMemcpyDeviceToHost, st));
        VDF_CUDA(ctx, cudaStreamSynchronize(st));
        if (left >= n_in) {
            // cannot happen: the smallest undecided vertex always has all smaller neighbours decided
            ctx->err = "greedy grouping made no progress";
            return VDF_ERR_CUDA;
        }
        wl_in = bufs[round & 1];
        n_in = left;
    }

    VDF_ALLOC(ctx, ctx->g_mk.ensure(nv * 8));
    VDF_ALLOC(ctx, ctx->g_mks.ensure(nv * 8));
    member_keys_kernel<<<blocks(nv), B, 0, st>>>(rks, state, ctx->g_parent.as<uint32_t>(), ctx->g_wl0.as<uint32_t>(), nv,
                                                 ctx->g_mk.as<uint64_t>(), cnt + 3);
    VDF_LAUNCHED(ctx);
    unsigned long long nm = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(&nm, cnt + 3, 8, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    return finish_groups(ctx, nm, out);
}

// ------------------------------------------------------------------------------------------------ connected components
// The north-star's "GPU union-find" as an OPTIONAL grouping mode (ctx option "grouping" = 1).
// It is NOT what the reference's search does (a chain a-b-c with a !~ c is one component but the greedy rule returns
// {a, b} only); it is what the app's DisjointSet (vid_dup_finder_app/src/app/disjoint_set.rs:22-44) computes for confirmed
// pairs.  Lock-free union-find: every edge hooks the larger root under the smaller one with atomicCAS (so the root of a
// component is its smallest vertex), finds use path halving.  Output convention is the greedy mode's: members ascending,
// the root last, groups by descending root -- for star-shaped clusters both modes return identical groups.
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t v) {
    uint32_t p = parent[v];
    while (p != v) {
        const uint32_t g = parent[p];
        if (g != p) parent[v] = g;  // path halving (benign race: only ever replaces a parent by an ancestor)
        v = p;
        p = g;
    }
    return v;
}

__global__ void uf_init_kernel(uint32_t* __restrict__ parent, uint64_t n) {
    uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v < n) parent[v] = (uint32_t)v;
}

__global__ void uf_union_kernel(const uint64_t* __restrict__ keys, uint64_t ne, uint32_t* parent) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    uint32_t a = uf_find(parent, (uint32_t)(keys[e] >> 32)), b = uf_find(parent, (uint32_t)keys[e]);
    while (a != b) {
        if (a > b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicCAS(&parent[b], b, a);  // hook root b under the smaller root a
        if (old == b) break;
        b = uf_find(parent, old);  // b was hooked by someone else meanwhile: continue from its new root
        a = uf_find(parent, a);
    }
}

__global__ void uf_member_keys_kernel(uint32_t* parent, uint64_t n, uint64_t* __restrict__ mk,
                                      unsigned long long* __restrict__ count) {
    uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t r = uf_find(parent, (uint32_t)v);
    if (r != (uint32_t)v) mk[atomicAdd(count, 1ull)] = ((uint64_t)(~r) << 32) | v;
}

int group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t ne, vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = nullptr;
    out->member_idx = nullptr;
    if (ne == 0 || n == 0) return empty_groups(out);
    const int B = 256;
    auto blocks = [&](uint64_t items) { return (unsigned)((items + B - 1) / B); };
    cudaStream_t st = ctx->stream;
    VDF_ALLOC(ctx, ctx->g_parent.ensure(n * 4));
    VDF_ALLOC(ctx, ctx->misc.ensure(64));
    unsigned long long* cnt = ctx->misc.as<unsigned long long>() + 4;
    VDF_CUDA(ctx, cudaMemsetAsync(cnt, 0, 32, st));
    uint32_t* parent = ctx->g_parent.as<uint32_t>();
    uf_init_kernel<<<blocks(n), B, 0, st>>>(parent, n);
    VDF_LAUNCHED(ctx);
    uf_union_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, parent);
    VDF_LAUNCHED(ctx);
    const uint64_t cap = std::min<uint64_t>(n, 2 * ne);  // every non-root vertex has an edge
    VDF_ALLOC(ctx, ctx->g_mk.ensure(cap * 8));
    VDF_ALLOC(ctx, ctx->g_mks.ensure(cap * 8));
    uf_member_keys_kernel<<<blocks(n), B, 0, st>>>(parent, n, ctx->g_mk.as<uint64_t>(), cnt + 3);
    VDF_LAUNCHED(ctx);
    unsigned long long nm = 0;
    VDF_CUDA(ctx, cudaMemcpyAsync(&nm, cnt + 3, 8, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    return finish_groups(ctx, nm, out);
}

}  // namespace vdf
