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
//
// Nothing here waits for the host between kernels: all rounds run inside ONE cooperative kernel (grid-wide barriers, the
// worklist counters live in HBM), and the CSR is built from the (i, j)-sorted edge list with two scans -- an edge (i, j) is
// a membership iff parent[j] == i, so the flagged edges ARE the groups, by ascending target with members ascending; the
// reference's descending order is index arithmetic.  The host reads {members, groups} once and then the CSR itself.
#include <cooperative_groups.h>

#include <algorithm>
#include <cub/device/device_scan.cuh>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace cg = cooperative_groups;

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

// cnt[0] = size of wl0 (seeds), cnt[1] / cnt[2] = sizes of the ping / pong worklists, cnt[3] = rounds run (diagnostic)
__global__ void __launch_bounds__(256) greedy_rounds_kernel(const uint64_t* __restrict__ rks, uint64_t ne, volatile uint8_t* state,
                                                            uint32_t* __restrict__ parent, const uint32_t* __restrict__ wl0,
                                                            uint32_t* __restrict__ wla, uint32_t* __restrict__ wlb,
                                                            unsigned long long* cnt) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t* wl_in = wl0;
    volatile unsigned long long* vcnt = cnt;
    unsigned long long n_in = vcnt[0];
    for (uint32_t round = 0; n_in > 0; ++round) {
        uint32_t* wl_out = (round & 1) ? wlb : wla;
        unsigned long long* n_out = cnt + 1 + (round & 1);
        for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n_in; k += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t e0 = wl_in[k];
            const uint32_t v = (uint32_t)(rks[e0] >> 32);
            uint8_t decided = kTarget;
            for (uint64_t e = e0; e < ne; ++e) {
                const uint64_t key = rks[e];
                if ((uint32_t)(key >> 32) != v) break;
                const uint32_t u = (uint32_t)key;
                const uint8_t s = state[u];
                if (s == kUndecided) {  // a smaller neighbour is still open: it may turn out to be the smallest target
                    wl_out[atomicAdd(n_out, 1ull)] = e0;
                    decided = kUndecided;
                    break;
                }
                if (s == kTarget) {  // everything before u is a member, so u is v's smallest target neighbour
                    parent[v] = u;
                    decided = kMember;
                    break;
                }
            }
            if (decided != kUndecided) {
                __threadfence();
                state[v] = decided;
            }
        }
        grid.sync();
        const unsigned long long left = vcnt[1 + (round & 1)];
        grid.sync();  // everybody has read `left` before the other counter is reset for the next round
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            vcnt[1 + ((round + 1) & 1)] = 0;
            vcnt[3] = round + 1;
        }
        grid.sync();
        // the smallest undecided vertex always has all smaller neighbours decided: left < n_in, the loop ends
        wl_in = wl_out;
        n_in = left;
    }
}

// keys = the (i << 32 | j)-sorted edge list: edge e is a membership iff j was consumed by target i
__global__ void member_flag_kernel(const uint64_t* __restrict__ keys, uint64_t ne, const uint8_t* __restrict__ state,
                                   const uint32_t* __restrict__ parent, uint32_t* __restrict__ flag) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const uint32_t i = (uint32_t)(keys[e] >> 32), j = (uint32_t)keys[e];
    flag[e] = (state[j] == kMember && parent[j] == i) ? 1u : 0u;
}

// compact the memberships: mk[rank - 1] = key; tot[0] = number of members
__global__ void member_compact_kernel(const uint64_t* __restrict__ keys, uint64_t ne, const uint32_t* __restrict__ flag,
                                      const uint32_t* __restrict__ rank, uint64_t* __restrict__ mk,
                                      unsigned long long* __restrict__ tot) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    if (flag[e]) mk[rank[e] - 1] = keys[e];
    if (e + 1 == ne) tot[0] = rank[e];
}

// gs[k] = 1 where member k opens a group (mk is sorted by (target, member)); zero beyond the members
__global__ void group_start_kernel(const uint64_t* __restrict__ mk, uint64_t ne, const unsigned long long* __restrict__ tot,
                                   uint32_t* __restrict__ gs) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= ne) return;
    const unsigned long long nm = tot[0];
    gs[k] = (k < nm && (k == 0 || (mk[k] >> 32) != (mk[k - 1] >> 32))) ? 1u : 0u;
}

// gnum[k] = 1-based ascending group number of member k; gstart[g] = first member of group g (1-based), gstart[ng + 1] = nm
__global__ void group_bounds_kernel(const uint32_t* __restrict__ gs, const uint32_t* __restrict__ gnum, uint64_t ne,
                                    unsigned long long* __restrict__ tot, uint32_t* __restrict__ gstart) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= ne) return;
    const unsigned long long nm = tot[0];
    if (k < nm && gs[k]) gstart[gnum[k]] = (uint32_t)k;
    if (k + 1 == ne) {
        const uint32_t ng = gnum[k];
        tot[1] = ng;
        gstart[ng + 1] = (uint32_t)nm;
    }
}

// CSR in the reference's order: group g (ascending target) becomes output group ng - g; inside a group the members
// ascending, then the target.  Ascending layout: group g occupies [gstart[g] + g - 1, gstart[g + 1] + g); the descending
// layout mirrors the blocks.  `remap` (optional) turns sorted positions into the caller's indices.
__global__ void group_csr_kernel(const uint64_t* __restrict__ mk, const uint32_t* __restrict__ gnum, const uint32_t* __restrict__ gstart,
                                 const unsigned long long* __restrict__ tot, uint64_t ne, const uint32_t* __restrict__ remap,
                                 uint64_t* __restrict__ group_ptr, uint64_t* __restrict__ member_idx) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= ne) return;
    const unsigned long long nm = tot[0], ng = tot[1];
    if (k >= nm) return;
    const uint32_t g = gnum[k];
    const uint64_t first = gstart[g], end_asc = (uint64_t)gstart[g + 1] + g;
    const uint64_t total = nm + ng, start = total - end_asc;
    const uint32_t member = (uint32_t)mk[k], target = (uint32_t)(mk[k] >> 32);
    member_idx[start + (k - first)] = remap ? remap[member] : member;
    if (k == first) {
        group_ptr[ng - g] = start;
        member_idx[start + (gstart[g + 1] - first)] = remap ? remap[target] : target;
        if (g == 1) group_ptr[ng] = total;
    }
}

// ---- result memory ---------------------------------------------------------------------------------------------------
// A vdf_groups is ONE block: group_ptr[n_groups + 1] followed by member_idx[].  The block comes from a small process-wide
// pool of pinned host buffers, so the device -> host copy of a result lands where the caller reads it (round 1 copied
// device -> pinned staging -> malloc'ed arrays: at 1 M hashes those copies were a third of a millisecond per search);
// vdf_free_groups hands the slot back.  When every slot is lent out the block is plain malloc memory.
namespace {
constexpr int kResultSlots = 8;
struct ResultSlot {
    void* p = nullptr;
    size_t cap = 0;
    bool busy = false;
};
ResultSlot g_slots[kResultSlots];
std::mutex g_slots_mu;
}  // namespace

void* result_alloc(size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_slots_mu);
        for (ResultSlot& sl : g_slots) {
            if (sl.busy) continue;
            if (sl.cap < bytes) {
                if (sl.p) cudaFreeHost(sl.p);
                sl.p = nullptr, sl.cap = 0;
                const size_t want = bytes + bytes / 4 + 4096;
                if (cudaMallocHost(&sl.p, want) != cudaSuccess) {
                    cudaGetLastError();
                    sl.p = nullptr;
                    break;
                }
                sl.cap = want;
            }
            sl.busy = true;
            return sl.p;
        }
    }
    return malloc(bytes ? bytes : 8);
}

void result_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_slots_mu);
        for (ResultSlot& sl : g_slots)
            if (sl.p == p) {
                sl.busy = false;
                return;
            }
    }
    free(p);
}

static int empty_groups(vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = (uint64_t*)result_alloc(2 * sizeof(uint64_t));
    if (!out->group_ptr) return VDF_ERR_ALLOC;
    out->group_ptr[0] = 0, out->group_ptr[1] = 0;
    out->member_idx = out->group_ptr + 1;
    return VDF_OK;
}

// tot[0] = members, tot[1] = groups are final on the stream; CSR in g_gp / g_mem -> host arrays
static int fetch_groups(vdf_ctx* ctx, const unsigned long long* d_tot, vdf_groups* out) {
    cudaStream_t st = ctx->stream;
    VDF_ALLOC(ctx, ctx->h_misc.ensure(256));
    unsigned long long* h = ctx->h_misc.as<unsigned long long>();
    VDF_CUDA(ctx, cudaMemcpyAsync(h, d_tot, 16, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    const uint64_t nm = h[0], ng = h[1];
    if (nm == 0) return empty_groups(out);
    const size_t gp_bytes = (size_t)(ng + 1) * 8, mem_bytes = (size_t)(nm + ng) * 8;
    out->n_groups = ng;
    out->group_ptr = (uint64_t*)result_alloc(gp_bytes + mem_bytes);  // pinned when a pool slot is free: the copies land in place
    if (!out->group_ptr) {
        ctx->err = "host allocation failed";
        return VDF_ERR_ALLOC;
    }
    out->member_idx = out->group_ptr + ng + 1;
    VDF_CUDA(ctx, cudaMemcpyAsync(out->group_ptr, ctx->g_gp.p, gp_bytes, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaMemcpyAsync(out->member_idx, ctx->g_mem.p, mem_bytes, cudaMemcpyDeviceToHost, st));
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->d2h += 16 + gp_bytes + mem_bytes;
    return VDF_OK;
}

static int scan_u32(vdf_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n) {
    size_t tmp = 0;
    VDF_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, (size_t)n, ctx->stream));
    VDF_ALLOC(ctx, ctx->sort_tmp.ensure(tmp));
    VDF_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->sort_tmp.p, tmp, in, out, (size_t)n, ctx->stream));
    ctx->launches += 1;
    return VDF_OK;
}

// mk = `count` keys (root << 32 | member), sorted by (root asc, member asc), count known only on the device (d_tot[0]) but
// bounded by `bound` -> CSR in g_gp / g_mem (descending root order, root last), then to the host
static int finish_groups(vdf_ctx* ctx, const uint64_t* mk, uint64_t bound, unsigned long long* d_tot, const uint32_t* d_remap,
                         vdf_groups* out) {
    const int B = 256;
    auto blocks = [&](uint64_t items) { return (unsigned)((items + B - 1) / B); };
    cudaStream_t st = ctx->stream;
    VDF_ALLOC(ctx, ctx->g_flag.ensure(bound * 4));
    VDF_ALLOC(ctx, ctx->g_scan.ensure(bound * 4));
    VDF_ALLOC(ctx, ctx->g_gstart.ensure((bound + 2) * 4));
    VDF_ALLOC(ctx, ctx->g_gp.ensure((bound + 1) * 8));
    VDF_ALLOC(ctx, ctx->g_mem.ensure(2 * bound * 8));
    group_start_kernel<<<blocks(bound), B, 0, st>>>(mk, bound, d_tot, ctx->g_flag.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    VDF_TRY(scan_u32(ctx, ctx->g_flag.as<uint32_t>(), ctx->g_scan.as<uint32_t>(), bound));
    group_bounds_kernel<<<blocks(bound), B, 0, st>>>(ctx->g_flag.as<uint32_t>(), ctx->g_scan.as<uint32_t>(), bound, d_tot,
                                                     ctx->g_gstart.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    group_csr_kernel<<<blocks(bound), B, 0, st>>>(mk, ctx->g_scan.as<uint32_t>(), ctx->g_gstart.as<uint32_t>(), d_tot, bound, d_remap,
                                                  ctx->g_gp.as<uint64_t>(), ctx->g_mem.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    return fetch_groups(ctx, d_tot, out);
}

int group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t ne, const uint32_t* d_remap, vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = nullptr;
    out->member_idx = nullptr;
    if (ne == 0 || n == 0) return empty_groups(out);
    if (ne > 0xFFFFFFFEull) {  // worklists and scans index edges with 32 bits
        ctx->err = "group_greedy: more than 2^32 - 2 edges";
        return VDF_ERR_INVALID;
    }
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
    VDF_ALLOC(ctx, ctx->g_cnt.ensure(64));
    unsigned long long* cnt = ctx->g_cnt.as<unsigned long long>();  // [0] seeds, [1] ping, [2] pong, [3] rounds, [4] members, [5] groups
    VDF_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));

    swap_halves_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, ctx->g_rk.as<uint64_t>());
    VDF_LAUNCHED(ctx);
    VDF_TRY(sort_keys(ctx, ctx->g_rk.as<uint64_t>(), ctx->g_rks.as<uint64_t>(), ne, 32 + bits_for(n)));  // (j << 32 | i), both < n
    const uint64_t* rks = ctx->g_rks.as<uint64_t>();
    uint8_t* state = ctx->g_state.as<uint8_t>();
    VDF_CUDA(ctx, cudaMemsetAsync(state, kTarget, n, st));
    seed_worklist_kernel<<<blocks(ne), B, 0, st>>>(rks, ne, state, ctx->g_wl0.as<uint32_t>(), cnt + 0);
    VDF_LAUNCHED(ctx);

    {   // all rounds in one cooperative launch: as many CTAs as are co-resident, at most what the seeds need
        if (ctx->greedy_blocks_per_sm == 0) {
            int per_sm = 0;
            VDF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_rounds_kernel, B, 0));
            ctx->greedy_blocks_per_sm = per_sm > 0 ? (per_sm > 4 ? 4 : per_sm) : 1;
        }
        unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->sm_count * ctx->greedy_blocks_per_sm, (uint64_t)blocks(ne));
        uint32_t* parent = ctx->g_parent.as<uint32_t>();
        const uint32_t* wl0 = ctx->g_wl0.as<uint32_t>();
        uint32_t* wla = ctx->g_wla.as<uint32_t>();
        uint32_t* wlb = ctx->g_wlb.as<uint32_t>();
        uint64_t ne_arg = ne;
        void* args[] = {(void*)&rks, (void*)&ne_arg, (void*)&state, (void*)&parent, (void*)&wl0, (void*)&wla, (void*)&wlb, (void*)&cnt};
        VDF_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)greedy_rounds_kernel, dim3(grid), dim3(B), args, 0, st));
        ctx->launches++;
    }

    VDF_ALLOC(ctx, ctx->g_flag.ensure(ne * 4));
    VDF_ALLOC(ctx, ctx->g_scan.ensure(ne * 4));
    VDF_ALLOC(ctx, ctx->g_mk.ensure(ne * 8));
    member_flag_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, state, ctx->g_parent.as<uint32_t>(), ctx->g_flag.as<uint32_t>());
    VDF_LAUNCHED(ctx);
    VDF_TRY(scan_u32(ctx, ctx->g_flag.as<uint32_t>(), ctx->g_scan.as<uint32_t>(), ne));
    member_compact_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, ctx->g_flag.as<uint32_t>(), ctx->g_scan.as<uint32_t>(),
                                                    ctx->g_mk.as<uint64_t>(), cnt + 4);
    VDF_LAUNCHED(ctx);
    return finish_groups(ctx, ctx->g_mk.as<uint64_t>(), ne, cnt + 4, d_remap, out);
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

// one key per non-root vertex: (root << 32 | vertex)
__global__ void uf_member_keys_kernel(uint32_t* parent, uint64_t n, uint64_t* __restrict__ mk, unsigned long long* __restrict__ count) {
    uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t r = uf_find(parent, (uint32_t)v);
    if (r != (uint32_t)v) mk[atomicAdd(count, 1ull)] = ((uint64_t)r << 32) | v;
}

int group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t ne, const uint32_t* d_remap, vdf_groups* out) {
    out->n_groups = 0;
    out->group_ptr = nullptr;
    out->member_idx = nullptr;
    if (ne == 0 || n == 0) return empty_groups(out);
    if (ne > 0xFFFFFFFEull) {
        ctx->err = "group_components: more than 2^32 - 2 edges";
        return VDF_ERR_INVALID;
    }
    const int B = 256;
    auto blocks = [&](uint64_t items) { return (unsigned)((items + B - 1) / B); };
    cudaStream_t st = ctx->stream;
    VDF_ALLOC(ctx, ctx->g_parent.ensure(n * 4));
    VDF_ALLOC(ctx, ctx->g_cnt.ensure(64));
    unsigned long long* cnt = ctx->g_cnt.as<unsigned long long>();
    VDF_CUDA(ctx, cudaMemsetAsync(cnt, 0, 64, st));
    uint32_t* parent = ctx->g_parent.as<uint32_t>();
    uf_init_kernel<<<blocks(n), B, 0, st>>>(parent, n);
    VDF_LAUNCHED(ctx);
    uf_union_kernel<<<blocks(ne), B, 0, st>>>(d_keys, ne, parent);
    VDF_LAUNCHED(ctx);
    const uint64_t cap = std::min<uint64_t>(n, 2 * ne);  // every non-root vertex has an edge
    VDF_ALLOC(ctx, ctx->g_mk.ensure(cap * 8));
    VDF_ALLOC(ctx, ctx->g_mks.ensure(cap * 8));
    // unfilled slots sort to the end and are ignored (the member count stays on the device)
    VDF_CUDA(ctx, cudaMemsetAsync(ctx->g_mk.p, 0xFF, cap * 8, st));
    uf_member_keys_kernel<<<blocks(n), B, 0, st>>>(parent, n, ctx->g_mk.as<uint64_t>(), cnt + 4);
    VDF_LAUNCHED(ctx);
    VDF_TRY(sort_keys(ctx, ctx->g_mk.as<uint64_t>(), ctx->g_mks.as<uint64_t>(), cap));
    return finish_groups(ctx, ctx->g_mks.as<uint64_t>(), cap, cnt + 4, d_remap, out);
}

}  // namespace vdf
