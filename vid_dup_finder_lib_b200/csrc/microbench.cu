// microbench.cu -- measures the per-SM instruction rates the rooflines in DESIGN.md are stated against:
// POPC (the search kernel's bound), LOP3 / IADD3 (what the carry-save variant trades POPC for), dp4a / dp2a and
// legacy IMMA (candidates for the exact-integer resize), plus a streaming read for the HBM figure.
// Usage: vdf_microbench  -> one JSON object per line on stdout.
#include <cublasLt.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 16;  // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(uint32_t* out, uint32_t seed) {
    uint32_t v[CH], a[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) v[k] = seed + threadIdx.x * 17 + k, a[k] = k;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (OP == 0) {  // POPC fed by a full-rate LOP3 and drained by an add: bound by the POPC pipe alone
                uint32_t x, pc;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x) : "r"(v[k]), "r"(a[k]), "r"(seed));
                asm volatile("popc.b32 %0, %1;" : "=r"(pc) : "r"(x));
                v[k] += pc;
            }
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(a[k]), "r"(seed));
            if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(a[k]));
            if (OP == 3) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(v[k]) : "r"(a[k]), "r"(seed));
            if (OP == 4) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[k]) : "r"(a[k]), "r"(seed));
            if (OP == 5) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(v[k]) : "r"(a[k]), "r"(seed));
            if (OP == 6) {  // XOR + POPC + add: the plain search inner step
                v[k] += __popc(a[k] ^ v[k]);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += v[k];
    if (s == 0xdeadbeef) out[0] = s;
}

// legacy tensor path: mma.sync m16n8k32 u8 x s8 -> s32
__global__ void __launch_bounds__(256) imma_kernel(int32_t* out, uint32_t seed) {
    int32_t c[4][4] = {};
    uint32_t a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, b0 = seed + 4, b1 = seed + 5;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[k][0]), "+r"(c[k][1]), "+r"(c[k][2]), "+r"(c[k][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int32_t s = 0;
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (s == 0x7eadbeef) out[0] = s;
}

__global__ void __launch_bounds__(256) stream_read_kernel(const uint4* __restrict__ in, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = in[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}


// ------------------------------------------------------------------------------------------------ tcgen05 issue rate
// The rate the tensor-core search kernels (search_tc.cu, variants 5 and 6) are bounded by: one CTA pair per two SMs issues
// back-to-back tcgen05.mma.cta_group::2 on resident shared-memory operands (random {0,1}-valued data in the same SWIZZLE_128B
// K-major layout and value encoding as the search kernel, so the power draw is comparable), two accumulators in flight,
// nothing else running.  KIND 0: kind::i8, M 256 x N 256 x K 32.  KIND 1: kind::mxf4 (block scales 1.0), M 256 x N 192 x K 64.
__device__ __forceinline__ uint32_t mb_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(mb_smem(bar)), "r"(parity)
                     : "memory");
}
__device__ __forceinline__ uint64_t mb_desc(uint32_t a) {  // K-major SWIZZLE_128B, SBO 1024 B (as search_tc.cu tc_desc)
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate_kernel(uint32_t iters, uint32_t seed, uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t mb_raw[];
    uint8_t* base = mb_raw + ((1024u - (mb_smem(mb_raw) & 1023u)) & 1023u);
    uint8_t* sA = base;            // 4 K-chunks x 128 rows x 128 B
    uint8_t* sB = base + 65536;    // 4 K-chunks x 128 rows x 128 B (kind::mxf4 reads 96 rows of each)
    uint64_t* done = reinterpret_cast<uint64_t*>(base + 131072);
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cr;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cr));
    uint32_t x = seed * 2654435761u + blockIdx.x * 40503u + tid * 9781u + 1u;
    for (int i = tid; i < 131072 / 4; i += 128) {  // random bits in the operand encodings of the search kernels
        x ^= x << 13, x ^= x >> 17, x ^= x << 5;
        reinterpret_cast<uint32_t*>(base)[i] = KIND == 0 ? (x & 0x01010101u) << (i & 7) : (x & 0x22222222u);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_smem(&done[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_smem(&done[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mb_smem(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tmem = *slot;
    if (KIND == 1) {  // UE8M0 1.0 in every byte of columns [384, 512)
        const uint32_t lanes = ((uint32_t)warp * 32) << 16;
        for (int c = 384; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + lanes + c), "r"(0x7F7F7F7Fu)
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0 && cr == 0) {
        const uint64_t da = mb_desc(mb_smem(sA)), db = mb_desc(mb_smem(sB));
        constexpr uint32_t idesc_i8 = (2u << 4) | (32u << 17) | (16u << 24);
        constexpr uint32_t idesc_f4 = (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | (1u << 23) | ((256u >> 4) << 24);
        constexpr uint32_t N = KIND == 0 ? 256 : 192;
        for (uint32_t it = 0; it < iters; ++it) {
            const uint32_t buf = it & 1;
            if (it >= 2) mb_wait(&done[buf], ((it >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pred;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
            if (pred) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint64_t off = (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                                         tmem + buf * N),
                                     "l"(da + off), "l"(db + off), "r"(idesc_i8), "r"((uint32_t)(k != 0))
                                     : "memory");
                    else
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(
                                tmem + buf * N),
                            "l"(da + off), "l"(db + off), "r"(idesc_f4), "r"((uint32_t)(k != 0)), "r"(tmem + 384), "r"(tmem + 448)
                            : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                                 mb_smem(&done[buf])),
                             "h"((uint16_t)1)
                             : "memory");
            }
            __syncwarp();
        }
        if (iters >= 2) mb_wait(&done[iters & 1], ((iters >> 1) - 1) & 1);  // batch iters - 2
        if (iters >= 1) mb_wait(&done[(iters - 1) & 1], (((iters - 1) >> 1)) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
    if (iters == 0xFFFFFFFFu) out[0] = tmem;
}

// tensor-memory read rate: every warp of the CTA streams tcgen05.ld.32x32b.x64 (8 KB per warp-load) over its lane quarter;
// this is what bounds the epilogue of the search kernels (one fp32 / s32 accumulator read per pair)
#define MB_R8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
__global__ void __launch_bounds__(256, 1) ldtm_rate_kernel(uint32_t iters, uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mb_smem(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot + (((uint32_t)(warp & 3) * 32) << 16);
    uint32_t sink = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t v[64];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,"
            "%62,%63}, [%64];"
            : MB_R8(v, 0), MB_R8(v, 8), MB_R8(v, 16), MB_R8(v, 24), MB_R8(v, 32), MB_R8(v, 40), MB_R8(v, 48), MB_R8(v, 56)
            : "r"(tmem + ((it & 7) * 64))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 64; k += 16) sink ^= v[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
    if (sink == 0x12345678u) out[0] = sink;
}

template <typename F>
static float time_ms(F&& launch, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

// An INDEPENDENT 4-bit tensor-core figure ("peak provenance"): cuBLASLt's block-scaled FP4 GEMM (NVFP4: e2m1
// operands, one UE4M3 scale per 16 values, all scales 1.0), 8192^3, bf16 output.  Two operand fillings -- random nibbles and
// the {0, 1.0} nibbles our kernel feeds the tensor cores -- as a burst (best of 10) and sustained for 4 s.  Same pipe as
// tcgen05.mma kind::mxf4 (kind::mxf4nvf4 issues at the same rate); a library GEMM also moves operands and writes a result, so
// it is a lower bound on the pipe, next to the bare issue-rate loop above.
static int cublaslt_fp4(int sms, int clk_khz) {
    const int64_t M = 8192, N = 8192, K = 8192;
    cublasLtHandle_t lt;
    if (cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) {
        printf("{\"op\": \"cublaslt_nvfp4_gemm_8192\", \"error\": \"cublasLtCreate\"}\n");
        return 0;
    }
    void *A, *B, *D, *sa, *sb, *ws;
    const size_t ws_bytes = 256u << 20;
    CK(cudaMalloc(&A, M * K / 2));
    CK(cudaMalloc(&B, N * K / 2));
    CK(cudaMalloc(&D, M * N * 2));
    CK(cudaMalloc(&sa, M * K / 16 + 4096));
    CK(cudaMalloc(&sb, N * K / 16 + 4096));
    CK(cudaMalloc(&ws, ws_bytes));
    CK(cudaMemset(sa, 0x38, M * K / 16 + 4096));  // UE4M3 1.0 in every scale slot: the tiled scale layout does not matter
    CK(cudaMemset(sb, 0x38, N * K / 16 + 4096));
    cublasLtMatmulDesc_t op;
    cublasLtMatrixLayout_t la, lb, lc;
    cublasLtMatmulPreference_t pref;
    const cublasOperation_t tr = CUBLAS_OP_T, nt = CUBLAS_OP_N;
    const cublasLtMatmulMatrixScale_t mode = CUBLASLT_MATMUL_MATRIX_SCALE_VEC16_UE4M3;
    bool ok = cublasLtMatmulDescCreate(&op, CUBLAS_COMPUTE_32F, CUDA_R_32F) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSA, &tr, sizeof tr) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_TRANSB, &nt, sizeof nt) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_A_SCALE_MODE, &mode, sizeof mode) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_B_SCALE_MODE, &mode, sizeof mode) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_A_SCALE_POINTER, &sa, sizeof sa) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_B_SCALE_POINTER, &sb, sizeof sb) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatrixLayoutCreate(&la, CUDA_R_4F_E2M1, K, M, K) == CUBLAS_STATUS_SUCCESS;  // A^T: K x M, K-major
    ok = ok && cublasLtMatrixLayoutCreate(&lb, CUDA_R_4F_E2M1, K, N, K) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatrixLayoutCreate(&lc, CUDA_R_16BF, M, N, M) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulPreferenceCreate(&pref) == CUBLAS_STATUS_SUCCESS;
    ok = ok && cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes, sizeof ws_bytes) == CUBLAS_STATUS_SUCCESS;
    cublasLtMatmulHeuristicResult_t heur;
    int found = 0;
    if (ok) ok = cublasLtMatmulAlgoGetHeuristic(lt, op, la, lb, lc, lc, pref, 1, &heur, &found) == CUBLAS_STATUS_SUCCESS && found > 0;
    if (!ok) {
        printf("{\"op\": \"cublaslt_nvfp4_gemm_8192\", \"error\": \"no FP4 block-scaled GEMM from cuBLASLt on this box\"}\n");
        return 0;
    }
    const float alpha = 1.0f, beta = 0.0f;
    auto gemm = [&]() { return cublasLtMatmul(lt, op, &alpha, A, la, B, lb, &beta, D, lc, D, lc, &heur.algo, ws, ws_bytes, 0); };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double flops = 2.0 * M * N * K;
    for (int fill = 0; fill < 2; ++fill) {
        std::vector<uint8_t> h((size_t)M * K / 2);
        uint64_t x = 0x9E3779B97F4A7C15ull;
        for (auto& b : h) {
            x ^= x << 13, x ^= x >> 7, x ^= x << 17;
            b = fill == 0 ? (uint8_t)(x >> 32) : (uint8_t)(((x >> 32) & 1 ? 0x02 : 0) | ((x >> 33) & 1 ? 0x20 : 0));
        }
        CK(cudaMemcpy(A, h.data(), h.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data() + 4096, h.size() - 4096, cudaMemcpyHostToDevice));
        if (gemm() != CUBLAS_STATUS_SUCCESS) {
            printf("{\"op\": \"cublaslt_nvfp4_gemm_8192\", \"error\": \"cublasLtMatmul failed\"}\n");
            return 0;
        }
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 10; ++r) {
            CK(cudaEventRecord(e0));
            gemm();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        const auto t0 = std::chrono::steady_clock::now();
        int reps = 0;
        CK(cudaEventRecord(e0));
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 4.0) {
            for (int r = 0; r < 20; ++r) gemm();
            reps += 20;
            CK(cudaDeviceSynchronize());
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const char* name = fill == 0 ? "random nibbles" : "{0, 1.0} nibbles";
        printf("{\"op\": \"cublaslt_nvfp4_gemm_8192\", \"data\": \"%s\", \"run\": \"burst\", \"ms\": %.4f, \"ops_per_s\": %.4e, "
               "\"macs_per_clk_per_sm_at_max_clock\": %.1f}\n",
               name, best, flops / (best * 1e-3), flops / 2 / (best * 1e-3) / ((double)clk_khz * 1e3) / sms);
        printf("{\"op\": \"cublaslt_nvfp4_gemm_8192\", \"data\": \"%s\", \"run\": \"sustained\", \"ms\": %.4f, \"gemms\": %d, \"ops_per_s\": %.4e, "
               "\"macs_per_clk_per_sm_at_max_clock\": %.1f}\n",
               name, ms / reps, reps, flops * reps / (ms * 1e-3), flops / 2 * reps / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms);
        fflush(stdout);
    }
    cudaFree(A), cudaFree(B), cudaFree(D), cudaFree(sa), cudaFree(sb), cudaFree(ws);
    cublasLtDestroy(lt);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, 64));
    const int blocks = sms * 8;
    const char* names[] = {"popc_add", "lop3", "iadd", "dp4a", "dp2a", "imad", "xor_popc_add"};
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", prop.name, sms, clk_khz / 1000);
    auto report = [&](const char* name, float ms, double ops_per_thread_iter) {
        double ops = (double)blocks * 256 * ITERS * ops_per_thread_iter;
        double per_s = ops / (ms * 1e-3);
        printf("{\"op\": \"%s\", \"ms\": %.3f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_max_clock\": %.2f}\n", name, ms,
               per_s, per_s / ((double)clk_khz * 1e3) / sms);
    };
    report(names[0], time_ms([&] { rate_kernel<0><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[1], time_ms([&] { rate_kernel<1><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[2], time_ms([&] { rate_kernel<2><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[3], time_ms([&] { rate_kernel<3><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[4], time_ms([&] { rate_kernel<4><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[5], time_ms([&] { rate_kernel<5><<<blocks, 256>>>(d_out, 1); }), CH);
    report(names[6], time_ms([&] { rate_kernel<6><<<blocks, 256>>>(d_out, 1); }), CH);
    {
        float ms = time_ms([&] { imma_kernel<<<blocks, 256>>>((int32_t*)d_out, 1); });
        double macs = (double)blocks * 8 * ITERS * 4 * (16.0 * 8 * 32);
        printf("{\"op\": \"imma_m16n8k32_u8s8\", \"ms\": %.3f, \"macs_per_s\": %.4e, \"macs_per_clk_per_sm_at_max_clock\": %.1f}\n",
               ms, macs / (ms * 1e-3), macs / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms);
    }
    for (int kind = 0; kind < 2; ++kind) {  // tcgen05 issue rate: burst (~10 ms) and sustained (~150 ms, the power-capped state)
        const size_t smem = 131072 + 1024 + 64;
        if (kind == 0) CK(cudaFuncSetAttribute(umma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CK(cudaFuncSetAttribute(umma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const double macs_per_batch = 16.0 * 256 * (kind == 0 ? 256.0 * 32 : 192.0 * 64);
        for (int len = 0; len < 2; ++len) {
            const uint32_t iters = (len == 0 ? 10000u : 150000u);
            const int pairs = sms / 2;
            float ms = time_ms([&] {
                if (kind == 0) umma_rate_kernel<0><<<2 * pairs, 128, smem>>>(iters, 7, d_out);
                else umma_rate_kernel<1><<<2 * pairs, 128, smem>>>(iters, 7, d_out);
            }, 3);
            CK(cudaGetLastError());
            const double macs = macs_per_batch * iters * pairs;
            printf("{\"op\": \"%s\", \"run\": \"%s\", \"ms\": %.3f, \"macs_per_s\": %.4e, \"ops_per_s\": %.4e, "
                   "\"macs_per_clk_per_sm_at_max_clock\": %.1f}\n",
                   kind == 0 ? "tcgen05_i8_2cta_m256n256k32" : "tcgen05_mxf4_2cta_m256n192k64", len == 0 ? "burst" : "sustained", ms,
                   macs / (ms * 1e-3), 2 * macs / (ms * 1e-3), macs / (ms * 1e-3) / ((double)clk_khz * 1e3) / (2 * pairs));
        }
    }
    for (int warps = 4; warps <= 8; warps += 4) {
        const uint32_t iters = 20000;
        float ms = time_ms([&] { ldtm_rate_kernel<<<sms, warps * 32>>>(iters, d_out); }, 3);
        CK(cudaGetLastError());
        const double bytes = (double)sms * warps * iters * 8192.0;
        printf("{\"op\": \"tcgen05_ld_32x32b_x64\", \"warps\": %d, \"ms\": %.3f, \"bytes_per_s\": %.4e, \"bytes_per_clk_per_sm_at_max_clock\": %.1f}\n",
               warps, ms, bytes / (ms * 1e-3), bytes / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms);
    }
    {
        size_t bytes = (size_t)8 << 30;
        uint4* buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        float ms = time_ms([&] { stream_read_kernel<<<sms * 16, 256>>>(buf, bytes / 16, d_out); });
        printf("{\"op\": \"hbm_stream_read\", \"run\": \"burst\", \"ms\": %.3f, \"gb_per_s\": %.1f}\n", ms, bytes / (ms * 1e-3) / 1e9);
        {   // sustained: the same launch back to back for ~1.2 s (the GPU reaches its power cap; bench.py's long hashing run does too)
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            const int reps = (int)(1200.0f / ms) + 1;
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) stream_read_kernel<<<sms * 16, 256>>>(buf, bytes / 16, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float tot = 0;
            CK(cudaEventElapsedTime(&tot, e0, e1));
            printf("{\"op\": \"hbm_stream_read\", \"run\": \"sustained\", \"launches\": %d, \"ms\": %.3f, \"gb_per_s\": %.1f}\n", reps, tot / reps,
                   bytes / (tot / reps * 1e-3) / 1e9);
            // and the last 20 % of a second such run alone (the steady state)
            const int head = reps * 4 / 5;
            for (int r = 0; r < head; ++r) stream_read_kernel<<<sms * 16, 256>>>(buf, bytes / 16, d_out);
            CK(cudaEventRecord(e0));
            for (int r = head; r < reps; ++r) stream_read_kernel<<<sms * 16, 256>>>(buf, bytes / 16, d_out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&tot, e0, e1));
            printf("{\"op\": \"hbm_stream_read\", \"run\": \"sustained, last fifth\", \"launches\": %d, \"ms\": %.3f, \"gb_per_s\": %.1f}\n", reps - head,
                   tot / (reps - head), bytes / (tot / (reps - head) * 1e-3) / 1e9);
        }
        cudaFree(buf);
    }
    CK(cudaDeviceSynchronize());
    if (cublaslt_fp4(sms, clk_khz)) return 1;
    return 0;
}
