// Read-pattern microbenchmark for the resize kernel's producer (hash.cu resize_mma_kernel): one CTA per 1920x1080 u8 frame,
// tiles of R rows x C bytes brought to shared memory through a ring of S stages, consumed by a token read of the tile.  What it
// answers: how much of the linear stream-read rate does each tile shape / copy mechanism keep?  (profiles/README.md, round 2)
//   mode 0: cp.async 16-byte copies (the kernel's mechanism)      mode 1: cp.async.bulk (one row segment per elected-lane copy)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s (%d)\n", #x, cudaGetErrorString(e_), __LINE__);   \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

constexpr int W = 1920, H = 1080;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t n) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// R rows x C bytes per stage, S stages, THREADS threads
template <int R, int C, int S, int THREADS>
__global__ void __launch_bounds__(THREADS) tile_read_cpasync(const uint8_t* __restrict__ frames, uint32_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kPitch = C + 16;
    constexpr int kStage = R * kPitch;
    constexpr int kPerRow = C / 16;                   // 16-byte copies per row
    constexpr int kRowsPerPass = THREADS / kPerRow;   // rows one pass of the CTA covers
    constexpr int kCopies = R / kRowsPerPass;
    static_assert(THREADS % kPerRow == 0 && R % kRowsPerPass == 0, "shape");
    const uint8_t* img = frames + (size_t)blockIdx.x * W * H;
    const int tid = threadIdx.x;
    const int c16 = (tid % kPerRow) * 16, r0 = tid / kPerRow;
    const int n_rb = (H + R - 1) / R, n_kc = (W + C - 1) / C, total = n_rb * n_kc;
    int p_it = 0, p_kc = 0, p_rb = 0, p_stage = 0;
    auto issue = [&]() {
        if (p_it < total) {
            uint8_t* st = smem + p_stage * kStage;
            const bool xok = p_kc * C + c16 < W;
#pragma unroll
            for (int i = 0; i < kCopies; ++i) {
                const int row = p_rb * R + r0 + i * kRowsPerPass;
                const bool ok = xok && row < H;
                cp_async16(st + (r0 + i * kRowsPerPass) * kPitch + c16, ok ? img + (size_t)row * W + p_kc * C + c16 : img, ok ? 16u : 0u);
            }
            ++p_it;
            if (++p_stage == S) p_stage = 0;
            if (++p_kc == n_kc) { p_kc = 0; ++p_rb; }
        }
        cp_commit();
    };
    for (int q = 0; q < S - 1; ++q) issue();
    uint32_t acc = 0;
    int c_stage = 0;
    for (int it = 0; it < total; ++it) {
        cp_wait<S - 2>();
        __syncthreads();
        issue();
        const uint8_t* st = smem + c_stage * kStage;
#pragma unroll
        for (int i = 0; i < kCopies; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(st + (r0 + i * kRowsPerPass) * kPitch + c16);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        if (++c_stage == S) c_stage = 0;
    }
    cp_wait<0>();
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

// ---- bulk-copy variant: every row segment of a tile is one cp.async.bulk (C bytes), issued by the lanes of warp 0; completion through
// one mbarrier per stage with the tile's byte count
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
            (uint32_t)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b))
                 : "memory");
}

template <int R, int C, int S, int THREADS>
__global__ void __launch_bounds__(THREADS) tile_read_bulk(const uint8_t* __restrict__ frames, uint32_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[S];
    constexpr int kPitch = C + 16;
    constexpr int kStage = R * kPitch;
    constexpr int kPerRow = C / 16;
    constexpr int kRowsPerPass = THREADS / kPerRow;
    constexpr int kCopies = R / kRowsPerPass;
    const uint8_t* img = frames + (size_t)blockIdx.x * W * H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c16 = (tid % kPerRow) * 16, r0 = tid / kPerRow;
    const int n_rb = (H + R - 1) / R, n_kc = (W + C - 1) / C, total = n_rb * n_kc;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int p_it = 0, p_kc = 0, p_rb = 0, p_stage = 0;
    auto issue = [&]() {  // warp 0 only
        if (p_it < total) {
            uint8_t* st = smem + p_stage * kStage;
            const int x0 = p_kc * C, wbytes = min(C, W - x0);
            const int rows = min(R, H - p_rb * R);
            if (lane == 0) mbar_expect(&full[p_stage], (uint32_t)(rows * wbytes));
            __syncwarp();
            for (int r = lane; r < rows; r += 32) bulk_g2s(st + r * kPitch, img + (size_t)(p_rb * R + r) * W + x0, (uint32_t)wbytes, &full[p_stage]);
            ++p_it;
            if (++p_stage == S) p_stage = 0;
            if (++p_kc == n_kc) { p_kc = 0; ++p_rb; }
        }
    };
    if (warp == 0)
        for (int q = 0; q < S - 1; ++q) issue();
    uint32_t acc = 0;
    int c_stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < total; ++it) {
        mbar_wait(&full[c_stage], phase);
        __syncthreads();  // everyone is past the previous tile: its stage may be refilled
        if (warp == 0) issue();
        const uint8_t* st = smem + c_stage * kStage;
#pragma unroll
        for (int i = 0; i < kCopies; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(st + (r0 + i * kRowsPerPass) * kPitch + c16);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        if (++c_stage == S) { c_stage = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

template <typename F>
static float time_ms(F&& f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

template <int R, int C, int S, int THREADS>
static void run(const uint8_t* buf, uint32_t* out, int n_frames, int mode) {
    constexpr int smem = S * R * (C + 16);
    auto k = mode == 0 ? tile_read_cpasync<R, C, S, THREADS> : tile_read_bulk<R, C, S, THREADS>;
    CK(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, smem));
    float ms = time_ms([&] { k<<<n_frames, THREADS, smem>>>(buf, out); });
    CK(cudaGetLastError());
    printf("{\"op\": \"tile_read\", \"mode\": \"%s\", \"rows\": %d, \"bytes_per_row\": %d, \"stages\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"gb_per_s\": %.1f}\n",
           mode == 0 ? "cp.async16" : "cp.async.bulk", R, C, S, THREADS, occ, ms, (double)n_frames * W * H / (ms * 1e-3) / 1e9);
    fflush(stdout);
}


// ---- persistent variant: one CTA per SM slot, frames claimed from a counter, the producer runs ahead ACROSS frames (the ring never
// drains between frames).  Emulates the planned fused hashing kernel's pixel stream.
template <int R, int C, int S, int THREADS, bool BULK = false>
__global__ void __launch_bounds__(THREADS) tile_read_persist(const uint8_t* __restrict__ frames, uint32_t* __restrict__ out, uint32_t n_frames,
                                                             uint32_t* __restrict__ counter, uint32_t extra_b) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_claim[2];
    __shared__ uint64_t s_cbar[S];
    constexpr int kPitch = C + 16;
    constexpr int kStage = R * kPitch;
    constexpr int kPerRow = C / 16;
    constexpr int kRowsPerPass = THREADS / kPerRow;
    constexpr int kCopies = R / kRowsPerPass;
    const int tid = threadIdx.x;
    const int c16 = (tid % kPerRow) * 16, r0 = tid / kPerRow;
    const int n_rb = (H + R - 1) / R, n_kc = (W + C - 1) / C, per_frame = n_rb * n_kc;
    if (BULK && tid == 0) {
        for (int q = 0; q < S; ++q) mbar_init(&s_cbar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) s_claim[0] = atomicAdd(counter, 1u);
    __syncthreads();
    uint32_t p_frame = s_claim[0], p_j = 0;  // producer: frame, ordinal of the frame in this CTA
    int p_kc = 0, p_rb = 0, p_stage = 0;
    if (tid == 0) s_claim[1] = atomicAdd(counter, 1u);
    auto issue = [&]() {
        if (p_frame < n_frames) {
            const uint8_t* img = frames + (size_t)p_frame * W * H;
            uint8_t* st = smem + p_stage * (kStage + (int)extra_b);
            const bool xok = p_kc * C + c16 < W;
#pragma unroll
            for (int i = 0; i < kCopies; ++i) {
                const int row = p_rb * R + r0 + i * kRowsPerPass;
                const bool ok = xok && row < H;
                cp_async16(st + (r0 + i * kRowsPerPass) * kPitch + c16, ok ? img + (size_t)row * W + p_kc * C + c16 : img, ok ? 16u : 0u);
            }
            // the coefficient fragments that ride along (L2-resident): extra_b bytes per stage
            if (BULK) {
                if (tid == 0 && extra_b) {
                    mbar_expect(&s_cbar[p_stage], extra_b);
                    bulk_g2s(st + kStage, frames + (size_t)p_kc * extra_b, extra_b, &s_cbar[p_stage]);
                }
            } else {
                for (uint32_t q = tid * 16; q < extra_b; q += THREADS * 16) cp_async16(st + kStage + q, frames + (size_t)p_kc * extra_b + q, 16u);
            }
            if (++p_stage == S) p_stage = 0;
            if (++p_kc == n_kc) {
                p_kc = 0;
                if (++p_rb == n_rb) {
                    p_rb = 0;
                    ++p_j;
                    p_frame = s_claim[p_j & 1];                                    // written >= one barrier ago
                    if (tid == 0) s_claim[(p_j + 1) & 1] = atomicAdd(counter, 1u);  // read next at the following switch
                }
            }
        }
        cp_commit();
    };
    for (int q = 0; q < S - 1; ++q) {
        issue();
        __syncthreads();
    }
    uint32_t acc = 0;
    int c_stage = 0;
    uint32_t c_it = 0, c_par = 0;
    // done when the producer has run out of frames and every tile it issued has been consumed (all of this is CTA-uniform)
    while (!(p_frame >= n_frames && c_it == p_j * (uint32_t)per_frame)) {
        cp_wait<S - 2>();
        if (BULK && extra_b) mbar_wait(&s_cbar[c_stage], c_par);
        __syncthreads();
        issue();
        const uint8_t* st = smem + c_stage * (kStage + (int)extra_b);
#pragma unroll
        for (int i = 0; i < kCopies; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(st + (r0 + i * kRowsPerPass) * kPitch + c16);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        if (BULK && extra_b) acc += *reinterpret_cast<const uint32_t*>(st + kStage + (tid * 16) % extra_b);
        if (++c_stage == S) c_stage = 0, c_par ^= 1u;
        ++c_it;
    }
    cp_wait<0>();
    if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

template <int R, int C, int S, int THREADS, bool BULK = false>
static void run_persist(const uint8_t* buf, uint32_t* out, int n_frames, uint32_t* counter, int ctas_per_sm, uint32_t extra_b) {
    const int smem = S * (R * (C + 16) + (int)extra_b);
    auto k = tile_read_persist<R, C, S, THREADS, BULK>;
    CK(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    if (occ < ctas_per_sm) { printf("{\"op\": \"tile_read_persist\", \"skip\": \"occupancy %d < %d\"}\n", occ, ctas_per_sm); return; }
    float ms = time_ms([&] {
        CK(cudaMemsetAsync(counter, 0, 4));
        k<<<sms * ctas_per_sm, THREADS, smem>>>(buf, out, (uint32_t)n_frames, counter, extra_b);
    });
    CK(cudaGetLastError());
    printf("{\"op\": \"tile_read_persist%s\", \"rows\": %d, \"bytes_per_row\": %d, \"stages\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"coef_bytes_per_stage\": %u, \"ms\": %.4f, \"gb_per_s\": %.1f}\n",
           BULK ? " coef by cp.async.bulk" : "", R, C, S, THREADS, ctas_per_sm, extra_b, ms, (double)n_frames * W * H / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

__global__ void fill_random(uint4* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = i * 0x9E3779B97F4A7C15ull + 12345;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        p[i] = make_uint4((uint32_t)z, (uint32_t)(z >> 32), (uint32_t)(z * 3), (uint32_t)((z * 5) >> 32));
    }
}

int main(int argc, char** argv) {
    const int n_frames = 4096;
    uint8_t* buf;
    uint32_t* out;
    CK(cudaMalloc(&buf, (size_t)n_frames * W * H));
    CK(cudaMemset(buf, 1, (size_t)n_frames * W * H));
    if (argc > 1 && argv[1][0] == 'r') {  // random content instead of a constant
        fill_random<<<148 * 8, 256>>>(reinterpret_cast<uint4*>(buf), (size_t)n_frames * W * H / 16);
        CK(cudaDeviceSynchronize());
        printf("{\"note\": \"random frame content\"}\n");
    }
    CK(cudaMalloc(&out, n_frames * 4));

    uint32_t* counter;
    CK(cudaMalloc(&counter, 4));
    for (uint32_t eb : {0u, 1u}) {
        run_persist<128, 128, 4, 128>(buf, out, n_frames, counter, 2, eb * 4096);
        run_persist<128, 128, 4, 128>(buf, out, n_frames, counter, 1, eb * 4096);
        run_persist<128, 256, 3, 128>(buf, out, n_frames, counter, 1, eb * 8192);
        run_persist<128, 256, 4, 128>(buf, out, n_frames, counter, 1, eb * 8192);
        run_persist<128, 256, 4, 256>(buf, out, n_frames, counter, 1, eb * 8192);
        run_persist<64, 512, 4, 128>(buf, out, n_frames, counter, 1, eb * 16384);
        run_persist<64, 256, 4, 128>(buf, out, n_frames, counter, 2, eb * 8192);
        run_persist<256, 128, 4, 256>(buf, out, n_frames, counter, 1, eb * 4096);
    }
    run_persist<128, 256, 4, 128, true>(buf, out, n_frames, counter, 1, 8192);
    run_persist<128, 256, 4, 128, true>(buf, out, n_frames, counter, 1, 8208);
    run_persist<128, 256, 4, 256, true>(buf, out, n_frames, counter, 1, 8208);
    run_persist<64, 256, 6, 128, true>(buf, out, n_frames, counter, 1, 8208);
    run_persist<64, 256, 4, 128, true>(buf, out, n_frames, counter, 2, 8208);
    run_persist<128, 128, 4, 128, true>(buf, out, n_frames, counter, 2, 4112);
    run_persist<128, 256, 4, 128, true>(buf, out, n_frames, counter, 1, 4112);
    run_persist<128, 256, 5, 128, true>(buf, out, n_frames, counter, 1, 8208);
    if (argc > 1) return 0;
    for (int mode = 0; mode < 2; ++mode) {
        run<128, 128, 4, 128>(buf, out, n_frames, mode);  // the kernel's shape
        run<128, 128, 3, 128>(buf, out, n_frames, mode);
        run<64, 256, 4, 128>(buf, out, n_frames, mode);
        run<32, 512, 4, 128>(buf, out, n_frames, mode);
        run<16, 1024, 4, 128>(buf, out, n_frames, mode);
        run<64, 256, 6, 128>(buf, out, n_frames, mode);
        run<32, 512, 6, 128>(buf, out, n_frames, mode);
        run<128, 128, 6, 128>(buf, out, n_frames, mode);
        run<128, 128, 4, 256>(buf, out, n_frames, mode);
        run<64, 128, 4, 128>(buf, out, n_frames, mode);   // 8 KB stages, more CTAs per SM
        run<64, 128, 8, 128>(buf, out, n_frames, mode);
    }
    return 0;
}
