// microbench.cu -- measures the per-SM instruction rates the rooflines in DESIGN.md are stated against:
// POPC (the search kernel's bound), LOP3 / IADD3 (what the carry-save variant trades POPC for), dp4a / dp2a and
// legacy IMMA (candidates for the exact-integer resize), plus a streaming read for the HBM figure.
// Usage: vdf_microbench  -> one JSON object per line on stdout.
#include <cuda_runtime.h>

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
    {
        size_t bytes = (size_t)8 << 30;
        uint4* buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        float ms = time_ms([&] { stream_read_kernel<<<sms * 16, 256>>>(buf, bytes / 16, d_out); });
        printf("{\"op\": \"hbm_stream_read\", \"ms\": %.3f, \"gb_per_s\": %.1f}\n", ms, bytes / (ms * 1e-3) / 1e9);
        cudaFree(buf);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
