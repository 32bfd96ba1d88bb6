// motion.cu -- `Cropdetect::Motion` on the GPU (second half).
//
// Replaces MotiondetectCrop::from_frames (vid_dup_finder_common/src/motioncrop/autocrop_frames.rs:36-218) with its helpers
// (darkest_frame.rs:19-111, frame_change.rs:15-132, utils.rs:8-131, crop.rs:32-199): the crop a stack gets when a
// video-in-a-frame (a dark picture with motion inside a bright, static surround) is to be cut out:
//   1. minimum / maximum over all 16 frames; if the stack uses neither 0 nor 255 its contrast is stretched to [0, 255]   (:52-113)
//   2. letterbox union over ALL 16 (stretched) frames; everything outside it reads as white                              (:123-150)
//   3. twice (the second time with the first crop's area whitened too, utils.rs:126-130), `from_frames_one` (:220-310):
//        dark   = min over the frames < 210                                            (darkest_frame.rs:19-70)
//        motion = sum over consecutive frames of |a - b| where >= 8, normalised to u16, to u8, Gaussian blur sigma 2,
//                 threshold 20, L-inf close 5                                          (frame_change.rs:15-132)
//        dark opened by min(h / 10, 10) when h > 100; its 8-connected regions that touch motion; the LARGEST of them
//        (ties: the last in raster order of first pixels); its bounding box, shrunk twice by one pixel a side
//   4. of the (up to two) crops: aspect ratio <= 3, area > 0.8 of the larger; the one with the smaller top; else the letterbox
//
// The arithmetic of the third-party pieces (imageproc stretch / threshold / morphology / labelling, image's blur and u16 -> u8)
// is restated in oracle/motioncrop_oracle.py from their published behaviour and is NOT pinned by the reference's tests beyond
// the letterbox / region / selection logic; this file follows that oracle bit for bit
// (tests/test_gpu_hashing.py::test_motion_crop_matches_oracle).  One full-resolution pass over the 16 frames per
// `from_frames_one` (HBM-bound, like the resize); everything after it works on one W x H mask per stack.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace vdf {

struct MoState {
    uint32_t mn, mx, stretch, pad0;
    uint32_t lb[4];        // letterbox union over the stretched frames: left, right, top, bottom
    uint32_t acc_mn, acc_mx;
    unsigned long long best;  // (pixel count << 32 | root pixel) of the chosen region, 0: none
    uint32_t bx0, by0, bx1, by1;
    uint32_t crop_valid[2];
    uint32_t crop[2][4];
};

struct MoGeom {  // where stack s of the sub-batch keeps its W x H planes
    size_t plane;  // pixels reserved per stack in every scratch buffer
};

__device__ __forceinline__ uint32_t mo_lut(uint32_t v, uint32_t mn, uint32_t mx, uint32_t stretch) {
    if (!stretch) return v;
    const uint32_t c = min(max(v, mn), mx);  // imageproc::contrast::stretch_contrast: u16 integer map, truncating
    return (c - mn) * 255u / (mx - mn);
}

__global__ void mo_init_kernel(MoState* __restrict__ st, uint32_t n) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    MoState z;
    memset(&z, 0, sizeof z);
    z.mn = 255;
    st[s] = z;
}

// step 1: minimum and maximum over every pixel of every frame
__global__ void __launch_bounds__(256) mo_minmax_kernel(const uint8_t* __restrict__ frames, const StackDev* __restrict__ stacks, MoState* __restrict__ st) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t mn = 255, mx = 0;
    if (px < (uint64_t)sd.width * sd.height) {
        const uint32_t y = (uint32_t)(px / sd.width), x = (uint32_t)(px % sd.width);
        const uint8_t* p = frames + sd.offset + (uint64_t)y * sd.pitch + x;
#pragma unroll 4
        for (int t = 0; t < VDF_DCT_SIZE; ++t) {
            const uint32_t v = __ldg(p + (uint64_t)t * sd.frame_stride);
            mn = min(mn, v), mx = max(mx, v);
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn), mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st[s].mn, mn);
        atomicMax(&st[s].mx, mx);
    }
}

// the stretch decision (autocrop_frames.rs:107-113) and the per-stack pixel table the letterbox scan reads through
__global__ void mo_prepare_kernel(MoState* __restrict__ st, uint8_t* __restrict__ luts) {
    const uint32_t s = blockIdx.x, v = threadIdx.x;
    const uint32_t mn = st[s].mn, mx = st[s].mx;
    const uint32_t stretch = (mx != 255 && mn != 0 && mn < mx) ? 1u : 0u;
    luts[(size_t)s * 256 + v] = (uint8_t)mo_lut(v, mn, mx, stretch);
    if (v == 0) st[s].stretch = stretch;
}

__global__ void mo_set_lb_kernel(MoState* __restrict__ st, const uint32_t* __restrict__ lb, uint32_t n) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    for (int k = 0; k < 4; ++k) st[s].lb[k] = lb[s * 4 + k];
}

__global__ void mo_begin_pass_kernel(MoState* __restrict__ st, uint32_t n) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    st[s].acc_mn = 0xFFFFFFFFu, st[s].acc_mx = 0, st[s].best = 0;
    st[s].bx0 = 0xFFFFFFFFu, st[s].by0 = 0xFFFFFFFFu, st[s].bx1 = 0, st[s].by1 = 0;
}

// step 3, the full-resolution pass: darkest pixel and thresholded frame differences over the 16 frames as from_frames_one
// sees them (stretched; white outside the letterbox; white inside the first crop in the second pass)
__global__ void __launch_bounds__(256) mo_accumulate_kernel(const uint8_t* __restrict__ frames, const StackDev* __restrict__ stacks, MoState* __restrict__ st,
                                                            uint32_t pass, size_t plane, uint8_t* __restrict__ dark, uint16_t* __restrict__ acc) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK) return;
    const MoState z = st[s];
    if (pass == 1 && !z.crop_valid[0]) return;
    const uint32_t W = sd.width, H = sd.height;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t a = 0xFFFFFFFFu, b = 0;
    if (px < (uint64_t)W * H) {
        const uint32_t y = (uint32_t)(px / W), x = (uint32_t)(px % W);
        bool white = !(x >= z.lb[0] && x < W - z.lb[1] && y >= z.lb[2] && y < H - z.lb[3]);
        if (pass == 1) white = white || (x >= z.crop[0][0] && x < W - z.crop[0][1] && y >= z.crop[0][2] && y < H - z.crop[0][3]);
        uint32_t darkest = 255, sum = 0;
        if (!white) {
            const uint8_t* p = frames + sd.offset + (uint64_t)y * sd.pitch + x;
            uint32_t prev = mo_lut(__ldg(p), z.mn, z.mx, z.stretch);
            darkest = prev;
#pragma unroll 3
            for (int t = 1; t < VDF_DCT_SIZE; ++t) {
                const uint32_t v = mo_lut(__ldg(p + (uint64_t)t * sd.frame_stride), z.mn, z.mx, z.stretch);
                const uint32_t d = v > prev ? v - prev : prev - v;
                if (d >= 8) sum += d;  // frame_change.rs: differences below 8 are noise
                darkest = min(darkest, v);
                prev = v;
            }
        }
        dark[s * plane + px] = darkest < 210 ? 255 : 0;  // darkest_frame.rs:19-70
        acc[s * plane + px] = (uint16_t)sum;              // 15 x 255 fits
        a = b = sum;
    }
    a = __reduce_min_sync(0xffffffffu, a), b = __reduce_max_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && a != 0xFFFFFFFFu) {
        atomicMin(&st[s].acc_mn, a);
        atomicMax(&st[s].acc_mx, b);
    }
}

// frame_change.rs:109-132 + image's u16 -> u8: (acc - min) * (65535 / (max - min)) in f64, truncated, then (c + 128) / 257
__global__ void __launch_bounds__(256) mo_normalize_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                           const uint16_t* __restrict__ acc, uint8_t* __restrict__ out) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)sd.width * sd.height) return;
    const uint32_t mn = st[s].acc_mn, mx = st[s].acc_mx;
    uint32_t v16 = 0;
    if (mn != mx) {  // equal: 65535 / 0 = inf, 0 * inf = NaN, NaN as u16 = 0
        const double scale = 65535.0 / (double)(mx - mn);
        double v = (double)(acc[s * plane + px] - mn) * scale;
        v = fmin(fmax(v, 0.0), 65535.0);
        v16 = (uint32_t)v;
    }
    out[s * plane + px] = (uint8_t)((v16 + 128u) / 257u);
}

// image::imageops::blur(sigma = 2): separable Gaussian, taps c-4 .. c+4 clamped to the image and re-normalised, f32 with
// separately rounded multiplies and adds.  The five tap values are the oracle's f32 results of
// 1 / (sqrt(2 pi) sigma) * exp(-d^2 / (2 sigma^2)), d = 0..4.
__device__ __forceinline__ float mo_tap(int d) {
    d = d < 0 ? -d : d;
    return __uint_as_float(d == 0 ? 0x3e4c422au : d == 1 ? 0x3e3441e8u : d == 2 ? 0x3df7c72du : d == 3 ? 0x3d84a042u : 0x3cdd25a2u);
}
template <bool kVertical>
__global__ void __launch_bounds__(256) mo_blur_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                      const uint8_t* __restrict__ src8, const float* __restrict__ src32, float* __restrict__ dst32,
                                                      uint8_t* __restrict__ dst8) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint32_t W = sd.width, H = sd.height;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)W * H) return;
    const int y = (int)(px / W), x = (int)(px % W);
    const int c = kVertical ? y : x, n = kVertical ? (int)H : (int)W;
    const int left = min(max(c - 4, 0), n - 1), right = min(max(c + 5, left + 1), n);
    float total = 0.0f;
    for (int i = left; i < right; ++i) total = __fadd_rn(total, mo_tap(i - c));
    float t = 0.0f;
    for (int i = left; i < right; ++i) {
        const size_t q = s * plane + (kVertical ? (size_t)i * W + x : (size_t)y * W + i);
        const float v = kVertical ? (float)src8[q] : src32[q];
        t = __fadd_rn(t, __fmul_rn(v, __fdiv_rn(mo_tap(i - c), total)));
    }
    if (kVertical) {
        dst32[s * plane + px] = t;
    } else {
        const double r = floor(fmin(fmax((double)t, 0.0), 255.0) + 0.5);
        dst8[s * plane + px] = (uint32_t)r > 20u ? 255 : 0;  // imageproc threshold(.., 20): p > t
    }
}

// one axis of an L-inf dilation (kMax) or erosion of a {0, 255} mask by k: any / all over the window clamped to the image
// (imageproc::morphology through the distance transform: the image border is not background)
template <bool kMax, bool kVertical>
__global__ void __launch_bounds__(256) mo_filter_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                        const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int k_fixed) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint32_t W = sd.width, H = sd.height;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)W * H) return;
    const int y = (int)(px / W), x = (int)(px % W);
    // k_fixed < 0: the dark mask's opening, by min(h / 10, 10), only when h > 100 (darkest_frame.rs:84-110)
    const int k = k_fixed >= 0 ? k_fixed : (H > 100 ? min((int)H / 10, 10) : 0);
    const int c = kVertical ? y : x, n = kVertical ? (int)H : (int)W;
    const int lo = max(c - k, 0), hi = min(c + k, n - 1);
    bool any = false, all = true;
    for (int i = lo; i <= hi; ++i) {
        const bool on = src[s * plane + (kVertical ? (size_t)i * W + x : (size_t)y * W + i)] != 0;
        any |= on, all &= on;
    }
    dst[s * plane + px] = (kMax ? any : all) ? 255 : 0;
}

// 8-connected regions of the opened dark mask by union-find over pixels: the root of a region is its first pixel in raster
// order, which is also the order imageproc numbers its labels in
__device__ __forceinline__ uint32_t mo_find(uint32_t* L, uint32_t v) {
    uint32_t p = L[v];
    while (p != v) {
        const uint32_t g = L[p];
        if (g != p) L[v] = g;
        v = p;
        p = g;
    }
    return v;
}
__device__ __forceinline__ void mo_union(uint32_t* L, uint32_t a, uint32_t b) {
    a = mo_find(L, a), b = mo_find(L, b);
    while (a != b) {
        if (a > b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicCAS(&L[b], b, a);
        if (old == b) break;
        b = mo_find(L, old);
        a = mo_find(L, a);
    }
}
__global__ void __launch_bounds__(256) mo_label_init_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                            const uint8_t* __restrict__ pp, uint32_t* __restrict__ label, uint32_t* __restrict__ count,
                                                            uint8_t* __restrict__ keep) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)sd.width * sd.height) return;
    label[s * plane + px] = (uint32_t)px;
    count[s * plane + px] = 0;
    keep[s * plane + px] = 0;
    (void)pp;
}
__global__ void __launch_bounds__(256) mo_label_union_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                             const uint8_t* __restrict__ pp, uint32_t* __restrict__ label) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint32_t W = sd.width, H = sd.height;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)W * H) return;
    const uint8_t* m = pp + s * plane;
    if (!m[px]) return;
    uint32_t* L = label + s * plane;
    const uint32_t y = (uint32_t)(px / W), x = (uint32_t)(px % W);
    if (x + 1 < W && m[px + 1]) mo_union(L, (uint32_t)px, (uint32_t)px + 1);
    if (y + 1 < H) {
        if (x > 0 && m[px + W - 1]) mo_union(L, (uint32_t)px, (uint32_t)(px + W - 1));
        if (m[px + W]) mo_union(L, (uint32_t)px, (uint32_t)(px + W));
        if (x + 1 < W && m[px + W + 1]) mo_union(L, (uint32_t)px, (uint32_t)(px + W + 1));
    }
}
// sizes of the regions, and which of them touch motion (regions_in_mask, utils.rs:32-42)
__global__ void __launch_bounds__(256) mo_label_count_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                             const uint8_t* __restrict__ pp, const uint8_t* __restrict__ motion, uint32_t* __restrict__ label,
                                                             uint32_t* __restrict__ count, uint8_t* __restrict__ keep) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)sd.width * sd.height) return;
    if (!pp[s * plane + px]) return;
    const uint32_t root = mo_find(label + s * plane, (uint32_t)px);
    label[s * plane + px] = root;
    atomicAdd(&count[s * plane + root], 1u);
    if (motion[s * plane + px] == 255) keep[s * plane + root] = 1;
}
// largest_region (utils.rs:56-70): the most pixels; Iterator::max_by keeps the LAST maximum = the later label = the larger root
__global__ void __launch_bounds__(256) mo_best_kernel(const StackDev* __restrict__ stacks, MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                      const uint8_t* __restrict__ pp, const uint32_t* __restrict__ label, const uint32_t* __restrict__ count,
                                                      const uint8_t* __restrict__ keep) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)sd.width * sd.height) return;
    if (!pp[s * plane + px] || label[s * plane + px] != (uint32_t)px || !keep[s * plane + px]) return;
    atomicMax(&st[s].best, ((unsigned long long)count[s * plane + px] << 32) | (unsigned long long)px);
}
__global__ void __launch_bounds__(256) mo_bbox_kernel(const StackDev* __restrict__ stacks, MoState* __restrict__ st, uint32_t pass, size_t plane,
                                                      const uint8_t* __restrict__ pp, const uint32_t* __restrict__ label) {
    const uint32_t s = blockIdx.y;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK || (pass == 1 && !st[s].crop_valid[0])) return;
    const uint64_t px = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (px >= (uint64_t)sd.width * sd.height) return;
    const unsigned long long best = st[s].best;
    if (!best || !pp[s * plane + px] || label[s * plane + px] != (uint32_t)best) return;
    const uint32_t y = (uint32_t)(px / sd.width), x = (uint32_t)(px % sd.width);
    atomicMin(&st[s].bx0, x), atomicMax(&st[s].bx1, x);
    atomicMin(&st[s].by0, y), atomicMax(&st[s].by1, y);
}
// autocrop_frames.rs:285-310: bounding box -> Crop::from_topleft_and_dims (crop.rs:32-50), shrunk twice (crop.rs:165-181)
__global__ void mo_end_pass_kernel(const StackDev* __restrict__ stacks, MoState* __restrict__ st, uint32_t pass, uint32_t n) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    MoState& z = st[s];
    z.crop_valid[pass] = 0;
    if (stacks[s].status != VDF_STACK_OK || (pass == 1 && !z.crop_valid[0]) || !z.best) return;
    const uint32_t W = stacks[s].width, H = stacks[s].height;
    const uint32_t bw = z.bx1 - z.bx0 + 1, bh = z.by1 - z.by0 + 1;
    uint32_t c[4] = {z.bx0, W - bw - z.bx0, z.by0, H - bh - z.by0};
    if (c[0] | c[1] | c[2] | c[3]) {
        uint32_t e[4] = {c[0], c[1], c[2], c[3]};
        bool ok = true;
        for (int r = 0; r < 2 && ok; ++r) {
            for (int k = 0; k < 4; ++k) e[k] += 1;
            ok = e[0] + e[1] < W && e[2] + e[3] < H;
        }
        if (ok)
            for (int k = 0; k < 4; ++k) c[k] = e[k];
    }
    for (int k = 0; k < 4; ++k) z.crop[pass][k] = c[k];
    z.crop_valid[pass] = 1;
}
// autocrop_frames.rs:152-218: of the crops found, aspect ratio <= 3 and area > 0.8 of the larger; the smaller top; else the letterbox
__global__ void mo_select_kernel(const StackDev* __restrict__ stacks, const MoState* __restrict__ st, uint32_t n, uint32_t* __restrict__ crop) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t out[4] = {0, 0, 0, 0};
    if (stacks[s].status == VDF_STACK_OK) {
        const MoState& z = st[s];
        const uint32_t W = stacks[s].width, H = stacks[s].height;
        for (int k = 0; k < 4; ++k) out[k] = z.lb[k];
        double largest = 0;
        for (int p = 0; p < 2; ++p)
            if (z.crop_valid[p]) largest = fmax(largest, (double)(W - z.crop[p][0] - z.crop[p][1]) * (double)(H - z.crop[p][2] - z.crop[p][3]));
        int chosen = -1;
        for (int p = 0; p < 2; ++p) {
            if (!z.crop_valid[p]) continue;
            const double cw = W - z.crop[p][0] - z.crop[p][1], ch = H - z.crop[p][2] - z.crop[p][3];
            const double ar = cw > ch ? cw / ch : ch / cw;
            if (ar <= 3.0 && cw * ch > largest * 0.8 && (chosen < 0 || z.crop[p][2] < z.crop[chosen][2])) chosen = p;  // min_by_key: first minimum
        }
        if (chosen >= 0)
            for (int k = 0; k < 4; ++k) out[k] = z.crop[chosen][k];
    }
    for (int k = 0; k < 4; ++k) crop[s * 4 + k] = out[k];
}

int motion_crop_device(vdf_ctx* ctx, const uint8_t* d_frames, const StackDev* d_sd, const StackDev* h_sd, uint32_t n, uint32_t* d_crop) {
    if (n == 0) return VDF_OK;
    cudaStream_t st = ctx->stream;
    size_t plane = 1;
    for (uint32_t s = 0; s < n; ++s)
        if (h_sd[s].status == VDF_STACK_OK) plane = std::max(plane, (size_t)h_sd[s].width * h_sd[s].height);
    plane = (plane + 63) / 64 * 64;
    // 17 bytes of scratch per pixel and stack: sub-batches of stacks that fit ~1.5 GB
    const uint32_t per = (uint32_t)std::max<size_t>(1, std::min<size_t>(n, ((size_t)1536 << 20) / (plane * 17)));
    VDF_ALLOC(ctx, ctx->m_state.ensure((size_t)n * sizeof(MoState)));
    VDF_ALLOC(ctx, ctx->m_lut.ensure((size_t)n * 256));
    VDF_ALLOC(ctx, ctx->m_sides.ensure((size_t)n * 16 * 4 * 4));
    VDF_ALLOC(ctx, ctx->m_acc.ensure((size_t)per * plane * 2));
    VDF_ALLOC(ctx, ctx->m_a.ensure((size_t)per * plane));
    VDF_ALLOC(ctx, ctx->m_b.ensure((size_t)per * plane));
    VDF_ALLOC(ctx, ctx->m_c.ensure((size_t)per * plane));
    VDF_ALLOC(ctx, ctx->m_f32.ensure((size_t)per * plane * 4));
    VDF_ALLOC(ctx, ctx->m_label.ensure((size_t)per * plane * 8));
    MoState* ms = ctx->m_state.as<MoState>();
    uint8_t* luts = ctx->m_lut.as<uint8_t>();
    const unsigned nb = (n + 127) / 128;
    mo_init_kernel<<<nb, 128, 0, st>>>(ms, n);
    VDF_LAUNCHED(ctx);
    const unsigned px_blocks = (unsigned)((plane + 255) / 256);
    mo_minmax_kernel<<<dim3(px_blocks, n), 256, 0, st>>>(d_frames, d_sd, ms);
    VDF_LAUNCHED(ctx);
    mo_prepare_kernel<<<n, 256, 0, st>>>(ms, luts);
    VDF_LAUNCHED(ctx);
    VDF_TRY(letterbox_all_frames(ctx, d_frames, d_sd, n, luts, ctx->m_sides.as<uint32_t>(), d_crop));
    mo_set_lb_kernel<<<nb, 128, 0, st>>>(ms, d_crop, n);
    VDF_LAUNCHED(ctx);
    uint16_t* acc = ctx->m_acc.as<uint16_t>();
    uint8_t *dark = ctx->m_a.as<uint8_t>(), *ma = ctx->m_b.as<uint8_t>(), *mb = ctx->m_c.as<uint8_t>();
    float* f32 = ctx->m_f32.as<float>();
    uint32_t* label = ctx->m_label.as<uint32_t>();
    uint32_t* count = label + (size_t)per * plane;
    for (uint32_t s0 = 0; s0 < n; s0 += per) {
        const uint32_t cnt = std::min(per, n - s0);
        const dim3 grid(px_blocks, cnt);
        const StackDev* sd = d_sd + s0;
        MoState* z = ms + s0;
        for (uint32_t pass = 0; pass < 2; ++pass) {
            mo_begin_pass_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(z, cnt);
            mo_accumulate_kernel<<<grid, 256, 0, st>>>(d_frames, sd, z, pass, plane, dark, acc);
            mo_normalize_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, acc, ma);
            mo_blur_kernel<true><<<grid, 256, 0, st>>>(sd, z, pass, plane, ma, nullptr, f32, nullptr);
            mo_blur_kernel<false><<<grid, 256, 0, st>>>(sd, z, pass, plane, nullptr, f32, nullptr, ma);  // + threshold 20
            // motion: L-inf close by 5 = dilate, then erode (ma -> mb -> ma -> mb -> ma)
            mo_filter_kernel<true, false><<<grid, 256, 0, st>>>(sd, z, pass, plane, ma, mb, 5);
            mo_filter_kernel<true, true><<<grid, 256, 0, st>>>(sd, z, pass, plane, mb, ma, 5);
            mo_filter_kernel<false, false><<<grid, 256, 0, st>>>(sd, z, pass, plane, ma, mb, 5);
            mo_filter_kernel<false, true><<<grid, 256, 0, st>>>(sd, z, pass, plane, mb, ma, 5);
            // dark: L-inf open by min(h / 10, 10) when h > 100 = erode, then dilate (dark -> mb -> dark -> mb -> dark; k = 0 copies)
            mo_filter_kernel<false, false><<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, mb, -1);
            mo_filter_kernel<false, true><<<grid, 256, 0, st>>>(sd, z, pass, plane, mb, dark, -1);
            mo_filter_kernel<true, false><<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, mb, -1);
            mo_filter_kernel<true, true><<<grid, 256, 0, st>>>(sd, z, pass, plane, mb, dark, -1);
            mo_label_init_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, label, count, mb);
            mo_label_union_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, label);
            mo_label_count_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, ma, label, count, mb);
            mo_best_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, label, count, mb);
            mo_bbox_kernel<<<grid, 256, 0, st>>>(sd, z, pass, plane, dark, label);
            mo_end_pass_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(sd, z, pass, cnt);
            ctx->launches += 19;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                ctx->err = std::string("motion crop: ") + cudaGetErrorString(e);
                return VDF_ERR_CUDA;
            }
        }
    }
    mo_select_kernel<<<nb, 128, 0, st>>>(d_sd, ms, n, d_crop);
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

}  // namespace vdf
