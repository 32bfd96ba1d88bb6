// hash.cu -- frame stack -> VideoHash on the GPU.
//
// Replaces the compute tail of gen_hash (video_hash_builder.rs:214-223):
//   H2  letterbox_{strip0,panels,tail}_kernel + crop_combine_kernel   <- cropdetect_letterbox / letterbox_crop / Crop::union
//                                                        (vid_dup_finder_common/src/video_frames_gray.rs:38-128,201-210,
//                                                         crop.rs:53-68)
//   H3  resize kernels                                <- crop_resize_buf (resize_gray.rs:11-54): fast_image_resize
//                                                        Lanczos3 u8 convolution, horizontal pass into a u8 temp,
//                                                        then vertical pass, i16 coefficients, i32 accumulate
//   H4  dct_pack_kernel                               <- Dct3d::from_images + dct_3d (dct_3d.rs:15-53,
//                                                        raw_dct_ops.rs:107-142): f64, split-radix DCT-II(16) x3
//   H5  (same kernel)                                 <- hash_bits + Lsb0 pack (dct_3d.rs:55-66, video_hash.rs:63-70)
//
// Everything integer is bit-exact by construction.  The DCT runs in f64 with explicit round-to-nearest
// multiplies/adds (no FMA contraction) in the same operation order as a split-radix butterfly, so exact zeros
// (static or mirror-symmetric content) stay exact zeros, as they do in the reference's rustdct butterflies.
//
// The i16 coefficient tables depend only on the cropped size, are generated on the host in f64 (libm sin, like
// the reference) and cached in HBM per size; a lookup table in HBM (size -> table) lets a small kernel turn the crops into
// resize jobs ON THE DEVICE, so a call runs letterbox -> jobs -> resize+DCT+pack without the host in between.  A size met
// for the first time is a "miss": the host builds its table after the pass and only the missed stacks run again.
// Kernels per call: ONE (hash_fused_kernel, below): persistent, warp-specialised; it scans frames 0 and 8 for bars, turns the crops
// into resize jobs, reads every pixel of the crop windows once and writes the 128-byte hashes.  The per-frame kernels it replaced
// (letterbox_*_kernel, resize_mma_kernel: context option hash_fused = 0) are kept for comparison and for Cropdetect::Motion's scan.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vdf {

// ================================================================================ H2: letterbox crop detect

// A "group" is the set of threads that runs one of the block-level device functions below: a whole thread block (BAR = 0,
// __syncthreads) or a warp-specialised part of one (BAR = named barrier id, NT threads; the fused kernel's helper warps).
template <int BAR, int NT>
__device__ __forceinline__ void group_sync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}
constexpr int kLbTol = 16;  // LetterboxColour::AnyColour(16), video_frames_gray.rs:206
constexpr int kColPanel = 32;
constexpr int kRowPanel = 32;

// Decide one strip from its 256-bin histogram (one warp).  mode = LAST maximum (Iterator::max_by_key,
// video_frames_gray.rs:82-87); letterbox iff count(|p - mode| <= tol) / len > 0.9 (:89-100), evaluated as
// 10*count > 9*len (equivalent for every len < 2^28, see tests/test_oracle_letterbox.py).
__device__ __forceinline__ bool strip_is_letterbox(const uint32_t* hist, uint32_t len, int lane) {
    uint32_t best_c = 0;
    int best_v = -1;
    for (int v = lane; v < 256; v += 32) {
        uint32_t c = hist[v];
        if (c > best_c || (c == best_c && v > best_v)) best_c = c, best_v = v;
    }
    for (int o = 16; o; o >>= 1) {
        uint32_t oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        int ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        if (oc > best_c || (oc == best_c && ov > best_v)) best_c = oc, best_v = ov;
    }
    const int lo = max(best_v - kLbTol, 0), hi = min(best_v + kLbTol, 255);
    uint32_t cnt = 0;
    for (int v = lo + lane; v <= hi; v += 32) cnt += hist[v];
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return 10ull * cnt > 9ull * len;
}

// The scan of one side of one frame is a walk from the edge inwards, strip by strip, to the first strip that is not letterbox
// (video_frames_gray.rs:38-128).  Its RESULT is "the index of the first such strip" = a minimum over strips, so the walk
// does not have to be serial:
//   letterbox_strip0_kernel   one CTA per (stack, frame, side): strip 0 alone.  Most sides have no bar: done, count 0.
//                             The others go onto a work list.
//   letterbox_panels_kernel   persistent CTAs over (listed side, panel p < kLbSpec) items, panel-major: the first kLbSpec panels
//                             of 32 strips of every listed side at once, atomicMin of the first non-letterbox strip found.
//                             Items whose side is already settled below their panel are skipped.
//   letterbox_tail_kernel     the listed sides still unsettled after kLbSpec * 32 strips (bars wider than 256 px) walk on
//                             serially; all counts are clamped to the side's length.
// Round 2's first version walked panel after panel inside one CTA per side (up to 8 dependent round trips to memory for a
// pillarboxed 1080p frame, 3.5 waves of CTAs): 105-120 us per 256 stacks for 0.6 % of the bytes.
//
// A strip whose values span <= tol is letterbox whatever its mode is (every pixel is within tol of every other one:
// count == len and 10 len > 9 len) -- exact, and it is what real bars look like.  So every panel is first looked at through
// its per-strip minimum and maximum (registers, byte-wise SIMD min/max, no shared-memory traffic); histograms are built
// only for the strips that are not decided that way -- in practice the one panel where the picture starts.
// kLut (Cropdetect::Motion, motion.cu): every pixel goes through a per-stack 256-entry table first (the contrast stretch of
// autocrop_frames.rs:107-113, monotone, so minima and maxima map through it) and all 16 frames are scanned (n_fr = 16, step 1)
// instead of frames 0 and 8 (n_fr = 2, step 8: step_by(8).take(8) over 16 frames).
constexpr uint32_t kLbSpec = 8;            // panels examined speculatively, side by side
constexpr uint32_t kLbNone = 0xFFFFFFFFu;  // no non-letterbox strip found yet

// Byte-wise minimum / maximum through the 16-bit SIMD min / max of sm_90+ (VIMNMX[3].U16x2): a word's even and odd bytes
// as two u16x2 values.  The byte-wise __vminu4 / __vmaxu4 are emulated on this architecture (six logic instructions each)
// and made the scan issue-bound: 9 M of its 25 M warp instructions were their LOP3s.
__device__ __forceinline__ uint32_t lb_even(uint32_t v) { return __byte_perm(v, 0u, 0x4240); }  // bytes 0, 2
__device__ __forceinline__ uint32_t lb_odd(uint32_t v) { return __byte_perm(v, 0u, 0x4341); }   // bytes 1, 3
struct LbRange4 {  // per byte lane of a word column: [even | odd] x [min | max]
    uint32_t mn_e = 0x00FF00FFu, mn_o = 0x00FF00FFu, mx_e = 0u, mx_o = 0u;
    __device__ __forceinline__ void add2(uint32_t a, uint32_t b) {
        const uint32_t ae = lb_even(a), ao = lb_odd(a), be = lb_even(b), bo = lb_odd(b);
        mn_e = __vimin3_u16x2(mn_e, ae, be), mn_o = __vimin3_u16x2(mn_o, ao, bo);
        mx_e = __vimax3_u16x2(mx_e, ae, be), mx_o = __vimax3_u16x2(mx_o, ao, bo);
    }
    __device__ __forceinline__ void merge(uint32_t xor_lane) {
        mn_e = __vminu2(mn_e, __shfl_xor_sync(0xffffffffu, mn_e, xor_lane)), mn_o = __vminu2(mn_o, __shfl_xor_sync(0xffffffffu, mn_o, xor_lane));
        mx_e = __vmaxu2(mx_e, __shfl_xor_sync(0xffffffffu, mx_e, xor_lane)), mx_o = __vmaxu2(mx_o, __shfl_xor_sync(0xffffffffu, mx_o, xor_lane));
    }
    __device__ __forceinline__ uint32_t mn(int q) const { return ((q & 1 ? mn_o : mn_e) >> (q & 2 ? 16 : 0)) & 0xFFFFu; }  // byte q of the word
    __device__ __forceinline__ uint32_t mx(int q) const { return ((q & 1 ? mx_o : mx_e) >> (q & 2 ? 16 : 0)) & 0xFFFFu; }
};
struct LbRange1 {  // over all bytes seen
    uint32_t mn2 = 0x00FF00FFu, mx2 = 0u;
    __device__ __forceinline__ void add(uint32_t v) {
        const uint32_t e = lb_even(v), o = lb_odd(v);
        mn2 = __vimin3_u16x2(mn2, e, o), mx2 = __vimax3_u16x2(mx2, e, o);
    }
    __device__ __forceinline__ uint32_t mn() const { return min(mn2 & 0xFFFFu, mn2 >> 16); }
    __device__ __forceinline__ uint32_t mx() const { return max(mx2 & 0xFFFFu, mx2 >> 16); }
};

struct LbShared {
    uint32_t hist[kColPanel * 257];  // column panels: one histogram per strip; row panels: 4 sub-histograms per warp
    uint32_t flags[kColPanel];       // out: strip k of the panel is letterbox
    uint32_t mn[kColPanel], mx[kColPanel];
    uint32_t need_hist, skip;
    uint8_t lut[256];
};

// flags[k] <- strip base + k of `side` is letterbox, for the 32 strips of one panel (256 threads, all of them; synchronised on return)
template <bool kLut, int NT = 256, int BAR = 0>
__device__ __forceinline__ void lb_panel_flags(LbShared& sh, const uint8_t* __restrict__ img, uint32_t W, uint32_t H, uint32_t P, uint32_t side,
                                               uint32_t base) {
    auto M = [&](uint32_t v) -> uint32_t { return kLut ? (uint32_t)sh.lut[v] : v; };
    constexpr int NW = NT / 32, RPP = NT / 8;  // warps in the group; rows one pass of 8-lane word rows covers
    const int tid = threadIdx.x % NT, lane = tid & 31, warp = tid >> 5;  // groups start at a multiple of NT threads
    const bool cols = side < 2;  // 0 left, 1 right, 2 top, 3 bottom
    const uint32_t limit = cols ? W : H, len = cols ? H : W;
    if (cols) {
        const uint32_t px0 = side == 0 ? base : W - 32 - base;  // first pixel column of a full 32-column panel
        // full, 4-byte aligned panel: 8 word loads cover a row of the panel, RPP rows per pass
        const bool fast = base + kColPanel <= W && ((reinterpret_cast<uintptr_t>(img) | P | px0) & 3) == 0;
        const uint32_t wq = tid & 7, r0 = tid >> 3;
        const uint32_t* col4 = reinterpret_cast<const uint32_t*>(img + px0) + wq;
        const uint32_t P4 = P >> 2;
        bool need_hist = true;
        if (fast) {  // pass A: value range of every column of the panel
            if (tid < kColPanel) sh.mn[tid] = 255u, sh.mx[tid] = 0u;
            if (tid == 0) sh.need_hist = 0;
            group_sync<BAR, NT>();
            // rows beyond the last one re-read the last one: harmless for a minimum / maximum, and every update is a full pair
            LbRange4 rg;
            for (uint32_t y0 = r0; y0 < H; y0 += RPP * 40) {  // 40 loads in flight per thread: 1280 rows (1080p) in one round trip at NT = 256
                uint32_t v[40];
#pragma unroll
                for (int u = 0; u < 40; ++u) v[u] = __ldg(col4 + (uint64_t)min(y0 + RPP * u, H - 1) * P4);
#pragma unroll
                for (int u = 0; u < 40; u += 2) rg.add2(v[u], v[u + 1]);
            }
            // the four row-threads of a warp that share a word column: lanes l, l^8, l^16, l^24
            rg.merge(8), rg.merge(16);
            if (lane < 8) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t c = 4 * wq + q, k = side == 0 ? c : 31 - c;  // strip index inside the panel
                    atomicMin(&sh.mn[k], rg.mn(q));
                    atomicMax(&sh.mx[k], rg.mx(q));
                }
            }
            group_sync<BAR, NT>();
            if (tid < kColPanel) {
                const bool nar = M(sh.mx[tid]) - M(sh.mn[tid]) <= (uint32_t)kLbTol;  // H >= 1: max >= min
                sh.flags[tid] = nar;
                if (!nar) sh.need_hist = 1;
            }
            group_sync<BAR, NT>();
            need_hist = sh.need_hist != 0;
            if (need_hist) {
                // Only the FIRST strip that is not letterbox matters, and a strip with a wide value range almost always is
                // picture: its histogram alone (256 threads on one column, data in L1 / L2 by now) settles the panel.  Only if
                // that strip turns out to be letterbox after all (a noisy bar) do all the undecided strips get histograms.
                const uint32_t wide = __ballot_sync(0xffffffffu, !sh.flags[lane]);  // pass A left a wide strip: wide != 0
                const uint32_t k0 = __ffs(wide) - 1;
                for (int q = tid; q < 256; q += NT) sh.hist[q] = 0;
                group_sync<BAR, NT>();
                const uint8_t* col = img + (side == 0 ? base + k0 : W - 1 - (base + k0));
                for (uint32_t y0 = tid; y0 < H; y0 += NT * 8) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = y0 + NT * u < H ? (uint32_t)__ldg(col + (uint64_t)(y0 + NT * u) * P) : 0x100u;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (v[u] < 0x100u) atomicAdd(&sh.hist[M(v[u])], 1u);
                }
                group_sync<BAR, NT>();
                if (warp == 0) {
                    const bool ok = strip_is_letterbox(sh.hist, len, lane);
                    if (lane == 0) sh.need_hist = ok ? 1u : 0u;  // flags[k0] stays 0 unless the full pass below says otherwise
                }
                group_sync<BAR, NT>();
                need_hist = sh.need_hist != 0;
            }
        }
        if (need_hist) {  // pass B: histograms of every undecided strip (the data of pass A comes from L1 / L2 this time)
            for (uint32_t q = tid; q < kColPanel * 257; q += NT) sh.hist[q] = 0;
            group_sync<BAR, NT>();
            if (fast) {
                for (uint32_t y0 = r0; y0 < H; y0 += RPP * 16) {
                    uint32_t v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t y = y0 + RPP * u;
                        v[u] = y < H ? __ldg(col4 + (uint64_t)y * P4) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (y0 + RPP * u < H) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t c = 4 * wq + q, k = side == 0 ? c : 31 - c;
                                if (!sh.flags[k]) atomicAdd(&sh.hist[k * 257 + M((v[u] >> (8 * q)) & 255u)], 1u);
                            }
                        }
                    }
                }
            } else {
                const uint32_t idx = base + lane;
                const bool act = idx < W;
                const uint32_t x = side == 0 ? idx : W - 1 - idx;
                // 16 independent loads in flight per lane before the (shared-memory) histogram updates
                for (uint32_t y0 = warp; y0 < H; y0 += NW * 16) {
                    uint32_t v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t y = y0 + NW * u;
                        v[u] = (act && y < H) ? (uint32_t)__ldg(img + (uint64_t)y * P + x) : 0xFFFFu;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (v[u] != 0xFFFFu) atomicAdd(&sh.hist[lane * 257 + M(v[u])], 1u);
                }
            }
            group_sync<BAR, NT>();
            for (uint32_t k = warp; k < kColPanel; k += NW) {
                if (fast && sh.flags[k]) continue;  // decided by its value range
                bool ok = false;
                if (base + k < limit) ok = strip_is_letterbox(sh.hist + k * 257, len, lane);
                if (lane == 0) sh.flags[k] = ok;
            }
            group_sync<BAR, NT>();
        }
    } else {
        // rows: warp w decides rows base + w + NW i (i = 0 .. 32 / NW - 1) on its own, in its own four sub-histograms (lane & 3) that
        // keep same-value lanes from piling onto one counter
        constexpr int kRowsPerWarp = kRowPanel / NW;
        uint32_t* h4 = sh.hist + (warp * 4) * 257;
        uint32_t* hs = h4 + (lane & 3) * 257;
        const uint32_t W4 = W >> 2;
        // first the value range of the rows of this warp, four rows' loads in flight together (rows of <= 2048 px, aligned):
        // a panel inside a bar costs one round trip to memory per four rows of a warp
        uint32_t decided = 0;  // bit i: row i is narrow
        if (((reinterpret_cast<uintptr_t>(img) | P) & 3) == 0 && W4 <= 32 * 16) {
            for (int i0 = 0; i0 < kRowsPerWarp; i0 += 4) {
                LbRange1 rg[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t idx = base + warp + NW * (i0 + i);
                    if (idx < H) {
                        const uint32_t y = side == 2 ? idx : H - 1 - idx;
                        const uint8_t* row = img + (uint64_t)y * P;
                        const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const uint32_t q = u * 32 + lane;
                            if (q < W4) rg[i].add(__ldg(row4 + q));
                        }
                        const uint32_t xt = (W4 << 2) + lane;
                        if (xt < W) rg[i].add((uint32_t)__ldg(row + xt) * 0x01010101u);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t a = __reduce_min_sync(0xffffffffu, rg[i].mn()), b = __reduce_max_sync(0xffffffffu, rg[i].mx());
                    if (base + warp + NW * (i0 + i) < H && M(b) - M(a) <= (uint32_t)kLbTol) decided |= 1u << (i0 + i);
                }
            }
        }
        // one row (warp-wide): its value range first when it fits the registers, the histogram only if that fails
        auto row_is_letterbox = [&](uint32_t idx) -> bool {
            const uint32_t y = side == 2 ? idx : H - 1 - idx;
            const uint8_t* row = img + (uint64_t)y * P;
            bool ok;
            if ((reinterpret_cast<uintptr_t>(row) & 3) == 0 && W4 <= 32 * 16) {
                // the whole row is in registers (<= 2048 px): its value range first, the histogram only if that fails
                const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
                uint32_t v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t q = u * 32 + lane;
                    v[u] = q < W4 ? __ldg(row4 + q) : 0u;
                }
                const uint32_t xt = (W4 << 2) + lane;  // the <= 3 pixels after the last whole word
                const uint32_t tail = xt < W ? (uint32_t)__ldg(row + xt) : 0x100u;
                LbRange1 rg;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (u * 32 + lane < W4) rg.add(v[u]);
                uint32_t mn = rg.mn(), mx = rg.mx();
                if (tail < 0x100u) mn = min(mn, tail), mx = max(mx, tail);
                mn = __reduce_min_sync(0xffffffffu, mn), mx = __reduce_max_sync(0xffffffffu, mx);
                ok = M(mx) - M(mn) <= (uint32_t)kLbTol;
                if (!ok) {
                    for (int q = lane; q < 4 * 257; q += 32) h4[q] = 0;
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        if (u * 32 + lane < W4) {
                            atomicAdd(&hs[M(v[u] & 255u)], 1u);
                            atomicAdd(&hs[M((v[u] >> 8) & 255u)], 1u);
                            atomicAdd(&hs[M((v[u] >> 16) & 255u)], 1u);
                            atomicAdd(&hs[M(v[u] >> 24)], 1u);
                        }
                    }
                    if (tail < 0x100u) atomicAdd(&hs[M(tail)], 1u);
                    __syncwarp();
                    for (int q = lane; q < 256; q += 32) h4[q] += h4[257 + q] + h4[514 + q] + h4[771 + q];
                    __syncwarp();
                    ok = strip_is_letterbox(h4, len, lane);
                }
            } else {  // long or unaligned rows: histogram straight away
                for (int q = lane; q < 4 * 257; q += 32) h4[q] = 0;
                __syncwarp();
                uint32_t x_done = 0;
                if ((reinterpret_cast<uintptr_t>(row) & 3) == 0) {  // up to 16 x 4 pixels per lane in flight
                    const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
                    for (uint32_t q0 = 0; q0 < W4; q0 += 32 * 16) {
                        uint32_t v[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const uint32_t q = q0 + u * 32 + lane;
                            v[u] = q < W4 ? __ldg(row4 + q) : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            if (q0 + u * 32 + lane < W4) {
                                atomicAdd(&hs[M(v[u] & 255u)], 1u);
                                atomicAdd(&hs[M((v[u] >> 8) & 255u)], 1u);
                                atomicAdd(&hs[M((v[u] >> 16) & 255u)], 1u);
                                atomicAdd(&hs[M(v[u] >> 24)], 1u);
                            }
                        }
                    }
                    x_done = W4 << 2;
                }
                for (uint32_t x0 = x_done; x0 < W; x0 += 32 * 16) {
                    uint32_t v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t x = x0 + u * 32 + lane;
                        v[u] = x < W ? (uint32_t)__ldg(row + x) : 0x100u;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (v[u] < 0x100u) atomicAdd(&hs[M(v[u])], 1u);
                }
                __syncwarp();
                for (int q = lane; q < 256; q += 32) h4[q] += h4[257 + q] + h4[514 + q] + h4[771 + q];
                __syncwarp();
                ok = strip_is_letterbox(h4, len, lane);
            }
            return ok;
        };
        // the rows the range test has settled; of the others only the FIRST can end the walk: it alone is evaluated (by the warp
        // that owns it) and, being wide, it almost always is picture.  Only if it is letterbox after all do the rest follow.
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)kRowsPerWarp; ++i)
            if (lane == 0) sh.flags[warp + NW * i] = (base + warp + NW * i < H && (decided & (1u << i))) ? 1u : 0u;
        if (tid == 0) sh.need_hist = 0;
        group_sync<BAR, NT>();
        const uint32_t open = __ballot_sync(0xffffffffu, !sh.flags[lane] && base + lane < H);
        if (open == 0) return;  // every row of the panel that exists is narrow
        const uint32_t k0 = __ffs(open) - 1;
        if (warp == (int)(k0 % NW)) {
            const bool ok = row_is_letterbox(base + k0);
            if (lane == 0) sh.flags[k0] = ok, sh.need_hist = ok ? 1u : 0u;
        }
        group_sync<BAR, NT>();
        if (!sh.need_hist) return;
        for (uint32_t i = 0; i < (uint32_t)kRowsPerWarp; ++i) {
            const uint32_t k = warp + NW * i, idx = base + k;
            if (k <= k0 || idx >= H || (decided & (1u << i))) continue;
            const bool ok = row_is_letterbox(idx);
            if (lane == 0) sh.flags[k] = ok;
            __syncwarp();
        }
        group_sync<BAR, NT>();
    }
}

// first non-letterbox strip among the panel's flags (thread 0 of a CTA whose lb_panel_flags has returned), or kLbNone
__device__ __forceinline__ uint32_t lb_first_stop(const LbShared& sh, uint32_t base, uint32_t limit) {
    for (uint32_t k = 0; k < kColPanel && base + k < limit; ++k)
        if (!sh.flags[k]) return base + k;
    return kLbNone;
}

// strip 0 of one side (group of NT threads, all of them): is it letterbox?  hist = 256 shared counters, s_flag = one shared word
template <bool kLut, int NT, int BAR>
__device__ __forceinline__ bool lb_strip0(uint32_t* hist, uint32_t* s_flag, const uint8_t* s_lut, const uint8_t* __restrict__ img, uint32_t W, uint32_t H,
                                          uint32_t P, uint32_t side) {
    const int tid = threadIdx.x % NT, lane = tid & 31, warp = tid >> 5;
    auto M = [&](uint32_t v) -> uint32_t { return kLut ? (uint32_t)s_lut[v] : v; };
    for (int q = tid; q < 256; q += NT) hist[q] = 0;
    group_sync<BAR, NT>();
    const bool cols = side < 2;
    if (cols) {
        const uint8_t* col = img + (side == 0 ? 0u : W - 1);
        for (uint32_t y0 = tid; y0 < H; y0 += NT * 8) {  // eight loads in flight per thread
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = y0 + NT * u < H ? (uint32_t)__ldg(col + (uint64_t)(y0 + NT * u) * P) : 0x100u;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (v[u] < 0x100u) atomicAdd(&hist[M(v[u])], 1u);
        }
    } else {
        const uint8_t* row = img + (uint64_t)(side == 2 ? 0u : H - 1) * P;
        uint32_t x_done = 0;
        if ((reinterpret_cast<uintptr_t>(row) & 3) == 0) {  // words, eight loads in flight per thread (one round trip for rows <= 4096 px at NT = 128)
            const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
            const uint32_t W4 = W >> 2;
            for (uint32_t q0 = tid; q0 < W4; q0 += NT * 8) {
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = q0 + NT * u < W4 ? __ldg(row4 + q0 + NT * u) : 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (q0 + NT * u < W4) {
                        atomicAdd(&hist[M(v[u] & 255u)], 1u);
                        atomicAdd(&hist[M((v[u] >> 8) & 255u)], 1u);
                        atomicAdd(&hist[M((v[u] >> 16) & 255u)], 1u);
                        atomicAdd(&hist[M(v[u] >> 24)], 1u);
                    }
            }
            x_done = W4 << 2;
        }
        for (uint32_t x = x_done + tid; x < W; x += NT) atomicAdd(&hist[M(__ldg(row + x))], 1u);
    }
    group_sync<BAR, NT>();
    if (warp == 0) {
        const bool ok = strip_is_letterbox(hist, cols ? H : W, lane);
        if (lane == 0) *s_flag = ok ? 1u : 0u;
    }
    group_sync<BAR, NT>();
    return *s_flag != 0;
}

// grid = n_stacks * n_fr * 4 sides; 128 threads.  sides[b] <- 0 (strip 0 is picture) or kLbNone + an entry in the work list.
template <bool kLut>
__global__ void __launch_bounds__(128) letterbox_strip0_kernel(const uint8_t* __restrict__ frames, const StackDev* __restrict__ stacks,
                                                               uint32_t* __restrict__ sides /* [n][n_fr][4] l,r,t,b */, uint32_t n_fr, uint32_t fr_step,
                                                               const uint8_t* __restrict__ luts /* [n][256] */, uint32_t* __restrict__ work,
                                                               uint32_t* __restrict__ n_work) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_flag;
    __shared__ uint8_t s_lut[kLut ? 256 : 4];
    const uint32_t b = blockIdx.x, side = b & 3, fr = (b >> 2) % n_fr, s = (b >> 2) / n_fr;
    const StackDev sd = stacks[s];
    if (sd.status != VDF_STACK_OK) return;
    const int tid = threadIdx.x;
    if (kLut) {
        for (int q = tid; q < 256; q += 128) s_lut[q] = luts[(size_t)s * 256 + q];
        __syncthreads();
    }
    const uint8_t* img = frames + sd.offset + (uint64_t)(fr * fr_step) * sd.frame_stride;
    const bool ok = lb_strip0<kLut, 128, 0>(hist, &s_flag, s_lut, img, sd.width, sd.height, sd.pitch, side);
    if (tid == 0) {
        sides[b] = ok ? kLbNone : 0u;
        if (ok) work[atomicAdd(n_work, 1u)] = b;
    }
}

// persistent: item = p * n_work + e for panel p < kLbSpec of listed side e; 256 threads
template <bool kLut>
__global__ void __launch_bounds__(256, 4) letterbox_panels_kernel(const uint8_t* __restrict__ frames, const StackDev* __restrict__ stacks,
                                                                 uint32_t* __restrict__ sides, uint32_t n_fr, uint32_t fr_step,
                                                                 const uint8_t* __restrict__ luts, const uint32_t* __restrict__ work,
                                                                 const uint32_t* __restrict__ n_work) {
    __shared__ LbShared sh;
    const uint32_t nw = *n_work;
    const int tid = threadIdx.x;
    for (uint32_t item = blockIdx.x; item < nw * kLbSpec; item += gridDim.x) {
        const uint32_t p = item / nw, b = work[item % nw];
        const uint32_t side = b & 3, fr = (b >> 2) % n_fr, s = (b >> 2) / n_fr;
        const StackDev sd = stacks[s];
        const uint32_t limit = side < 2 ? sd.width : sd.height, base = p * kColPanel;
        __syncthreads();  // the previous item's flags have been read
        if (tid == 0) sh.skip = (base >= limit || __ldcg(&sides[b]) < base) ? 1u : 0u;  // settled nearer the edge: nothing to add
        if (kLut) sh.lut[tid] = luts[(size_t)s * 256 + tid];
        __syncthreads();
        if (sh.skip) continue;
        const uint8_t* img = frames + sd.offset + (uint64_t)(fr * fr_step) * sd.frame_stride;
        lb_panel_flags<kLut>(sh, img, sd.width, sd.height, sd.pitch, side, base);
        if (tid == 0) {
            const uint32_t f = lb_first_stop(sh, base, limit);
            if (f != kLbNone) atomicMin(&sides[b], f);
        }
    }
}

// persistent over the listed sides: the rare serial continuation beyond kLbSpec panels, and the clamp to the side's length
// (a side that is letterbox all the way counts every strip, video_frames_gray.rs:103-117)
template <bool kLut>
__global__ void __launch_bounds__(256, 4) letterbox_tail_kernel(const uint8_t* __restrict__ frames, const StackDev* __restrict__ stacks,
                                                               uint32_t* __restrict__ sides, uint32_t n_fr, uint32_t fr_step,
                                                               const uint8_t* __restrict__ luts, const uint32_t* __restrict__ work,
                                                               const uint32_t* __restrict__ n_work) {
    __shared__ LbShared sh;
    __shared__ uint32_t s_first;
    const uint32_t nw = *n_work;
    const int tid = threadIdx.x;
    for (uint32_t e = blockIdx.x; e < nw; e += gridDim.x) {
        const uint32_t b = work[e];
        const uint32_t side = b & 3, fr = (b >> 2) % n_fr, s = (b >> 2) / n_fr;
        const StackDev sd = stacks[s];
        const uint32_t limit = side < 2 ? sd.width : sd.height;
        __syncthreads();
        if (tid == 0) s_first = __ldcg(&sides[b]);
        if (kLut) sh.lut[tid] = luts[(size_t)s * 256 + tid];
        __syncthreads();
        if (s_first == kLbNone) {
            const uint8_t* img = frames + sd.offset + (uint64_t)(fr * fr_step) * sd.frame_stride;
            for (uint32_t base = kLbSpec * kColPanel; base < limit; base += kColPanel) {
                lb_panel_flags<kLut>(sh, img, sd.width, sd.height, sd.pitch, side, base);
                if (tid == 0) s_first = lb_first_stop(sh, base, limit);
                __syncthreads();
                if (s_first != kLbNone) break;
            }
        }
        if (tid == 0) sides[b] = min(s_first, limit);
    }
}

// per frame: keep (l,r,t,b) only if at least one pixel remains each way (video_frames_gray.rs:119-127);
// across frames 0 and 8: per-side minimum (Crop::union, crop.rs:53-68)
__global__ void crop_combine_kernel(const StackDev* __restrict__ stacks, const uint32_t* __restrict__ sides, uint32_t n, uint32_t n_fr,
                                    uint32_t* __restrict__ crop /* [n][4] */) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t out[4] = {0, 0, 0, 0};
    if (stacks[s].status == VDF_STACK_OK) {
        const int W = (int)stacks[s].width, H = (int)stacks[s].height;
        for (uint32_t fr = 0; fr < n_fr; ++fr) {
            uint32_t c[4];
            for (int k = 0; k < 4; ++k) c[k] = sides[((size_t)s * n_fr + fr) * 4 + k];
            if (!(W - (int)c[0] - (int)c[1] >= 1 && H - (int)c[2] - (int)c[3] >= 1)) c[0] = c[1] = c[2] = c[3] = 0;
            for (int k = 0; k < 4; ++k) out[k] = fr == 0 ? c[k] : min(out[k], c[k]);
        }
    }
    for (int k = 0; k < 4; ++k) crop[s * 4 + k] = out[k];
}

// ================================================================================ H4 + H5: 3-D DCT, threshold, pack
// twiddles: [0..7] n=16 (re,im) x4, [8..11] n=8 (re,im) x2, [12..13] n=4, [14] FRAC_1_SQRT_2
__constant__ double c_tw[16];

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ void dct2_2(double& b0, double& b1) {
    const double s = dadd(b0, b1);
    b1 = dmul(dsub(b0, b1), c_tw[14]);
    b0 = s;
}
__device__ __forceinline__ void dct2_4(double* b) {
    const double re = c_tw[12], im = c_tw[13];
    const double lower = dsub(b[0], b[3]), upper = dsub(b[2], b[1]);
    double e0 = dadd(b[0], b[3]), e1 = dadd(b[1], b[2]);
    dct2_2(e0, e1);
    b[0] = e0;
    b[1] = dsub(dmul(lower, re), dmul(upper, im));
    b[2] = e1;
    b[3] = dadd(dmul(upper, re), dmul(lower, im));
}
// split radix: one half-size DCT-II on the mirrored sums, two quarter-size DCT-IIs on the rotated differences
template <int N>
__device__ __forceinline__ void dct2_sr(double* x);
template <>
__device__ __forceinline__ void dct2_sr<2>(double* x) {
    dct2_2(x[0], x[1]);
}
template <>
__device__ __forceinline__ void dct2_sr<4>(double* x) {
    dct2_4(x);
}
template <int N>
__device__ __forceinline__ void dct2_sr(double* x) {
    constexpr int H = N / 2, Q = N / 4;
    constexpr int TW = (N == 16) ? 0 : 8;
    double d2[H], ev[Q], od[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const double bot = x[i], top = x[N - 1 - i];
        const double hb = x[H - 1 - i], ht = x[H + i];
        d2[i] = dadd(top, bot);
        d2[H - 1 - i] = dadd(hb, ht);
        const double lower = dsub(bot, top), upper = dsub(hb, ht);
        const double re = c_tw[TW + 2 * i], im = c_tw[TW + 2 * i + 1];
        const double c = dadd(dmul(lower, re), dmul(upper, im));
        const double s = dsub(dmul(upper, re), dmul(lower, im));
        ev[i] = c;
        od[Q - 1 - i] = (i % 2 == 0) ? s : -s;
    }
    dct2_sr<H>(d2);
    dct2_sr<Q>(ev);
    dct2_sr<Q>(od);
    x[0] = d2[0];
    x[1] = ev[0];
    x[2] = d2[1];
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        const double c = ev[i];
        const double s = ((i + Q) % 2 == 0) ? -od[Q - i] : od[Q - i];
        x[4 * i - 1] = dadd(c, s);
        x[4 * i] = d2[2 * i];
        x[4 * i + 1] = dsub(c, s);
        x[4 * i + 2] = d2[2 * i + 1];
    }
    x[N - 1] = -od[0];
}

// cube index with a 17-double row pitch: conflict-free for all three passes
__device__ __forceinline__ int cidx(int t, int x, int y) { return (t * 16 + x) * 17 + y; }

constexpr int kCubeBytes = 16 * 16 * 17 * 8;  // the f64 cube with its padded pitch: 34 816 B

// One thread block (NT = 128 or 256 threads, all of them): small [t][row][col] u8 -> m[t][x=col][y=row] = p - 128
// (dct_3d.rs:40-44,76), DCT along y, x, t (raw_dct_ops.rs:107-142), bit t*100+x*10+y = coef > 0.0 (dct_3d.rs:55-66), Lsb0
// words.  `cube` = kCubeBytes of shared memory.
template <int NT, int BAR = 0>
__device__ __forceinline__ void dct_pack_block(const uint8_t* __restrict__ sm, double* cube, uint32_t* __restrict__ out, int tid) {
    for (int q = tid; q < 4096; q += NT) {
        const int t = q >> 8, row = (q >> 4) & 15, col = q & 15;
        cube[cidx(t, col, row)] = (double)__ldcg(sm + q) - 128.0;  // written by other thread blocks: read through L2
    }
    group_sync<BAR, NT>();
    double v[16];
    for (int line = tid; line < 256; line += NT) {  // along y: line (t, x)
        const int t = line >> 4, x = line & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cube[cidx(t, x, k)];
        dct2_sr<16>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) cube[cidx(t, x, k)] = v[k];
    }
    group_sync<BAR, NT>();
    for (int line = tid; line < 256; line += NT) {  // along x: line (t, y)
        const int t = line >> 4, y = line & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cube[cidx(t, k, y)];
        dct2_sr<16>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) cube[cidx(t, k, y)] = v[k];
    }
    group_sync<BAR, NT>();
    for (int line = tid; line < 256; line += NT) {  // along t: line (x, y)
        const int x = line >> 4, y = line & 15;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cube[cidx(k, x, y)];
        dct2_sr<16>(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) cube[cidx(k, x, y)] = v[k];
    }
    group_sync<BAR, NT>();
    for (int b = tid; b < 1024; b += NT) {
        bool bit = false;
        if (b < VDF_HASH_BITS) {
            const int t = b / 100, x = (b / 10) % 10, y = b % 10;
            bit = cube[cidx(t, x, y)] > 0.0;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if ((tid & 31) == 0) out[b >> 5] = word;
    }
}

// stand-alone form (Dct3d::from_images + hash_bits on an already resized cube: vdf_hash_from_small, parity taps)
__global__ void __launch_bounds__(256) dct_pack_kernel(const uint8_t* __restrict__ small, const int32_t* __restrict__ status,
                                                       uint32_t n, uint32_t* __restrict__ out_hash /* [n][32] u32 */) {
    __shared__ double cube[16 * 16 * 17];
    const uint32_t s = blockIdx.x;
    const int tid = threadIdx.x;
    uint32_t* out = out_hash + (uint64_t)s * 32;
    if (status && status[s] != VDF_STACK_OK) {
        if (tid < 32) out[tid] = 0;
        return;
    }
    dct_pack_block<256>(small + (uint64_t)s * 4096, cube, out, tid);
}

struct StackJob;
__global__ void dct_pack_jobs_kernel(const uint8_t* __restrict__ small, const StackJob* __restrict__ jobs, uint32_t n, uint32_t* __restrict__ out_hash);

// The thread block that completes a stack's 16th frame turns the stack's cube (4 KB, just written, L2-resident) into the
// hash: every block publishes its 256 output bytes (fence), then counts itself in; the one that counts 16 hashes.
template <int NT>
__device__ __forceinline__ void finish_stack(uint8_t* __restrict__ small, uint32_t s, uint32_t* __restrict__ done, uint32_t* __restrict__ out_hash,
                                             double* cube, int tid) {
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&done[s], 1u) == 15u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    dct_pack_block<NT>(small + (uint64_t)s * 4096, cube, out_hash + (uint64_t)s * 32, tid);
}


// ================================================================================ H3: crop + Lanczos3 resize
constexpr int32_t kJobSkip = 100;  // internal status: nothing to do for this stack in this pass (missed table, or done already)

struct StackJob {
    uint64_t offset, frame_stride;
    uint32_t pitch, left, top, cw, ch;
    int32_t status;
    const uint32_t* bh;  // 16 x (start,size), horizontal
    const int16_t* kh;   // 16 x win_h
    const uint32_t* bv;
    const int16_t* kv;
    uint32_t win_h, prec_h, win_v, prec_v;
    // tensor-core path (resize_mma_kernel): present iff the stack's addresses are 16-byte aligned
    const uint2* kb;     // horizontal coefficients as IMMA B fragments: [k-step][n-tile 0..3][lane] (b0, b1)
    const uint8_t* kmask;  // per k-chunk: bit (2*ks + octet) set iff outputs 8*octet..+7 have a tap in k-step ks
    uint32_t x0_al;      // first column loaded = left rounded down to 16
    uint32_t n_kch;      // 128-pixel k-chunks covering [x0_al, left + cw)
    uint32_t fast;       // 1: resize_mma_kernel / hash_fused_kernel, 0: resize_general_kernel
    // fused kernel (hash_fused_kernel)
    const uint8_t* kb2;  // per 256-pixel k-chunk: the B fragments of its two 128-pixel halves (8 KB) + a 16-byte header {mask, mask}
    const uint4* kva;    // vertical coefficients as IMMA A fragments: [32-row k-step][hi, lo][lane]
    uint32_t n_kc2;      // 256-pixel k-chunks covering [x0_al, left + cw)
    uint32_t pad2;
};

__device__ __forceinline__ uint8_t clip8(int32_t v, uint32_t precision) {
    return (uint8_t)min(max(v >> precision, 0), 255);
}

// General path (any shape / alignment): one CTA per (stack, frame).  Horizontal pass: one thread per
// (row, output) walks its coefficient window; the u8 intermediate [ch][16] lives in shared memory
// (the reference rounds to u8 between the passes); vertical pass: one thread per output pixel.
__global__ void __launch_bounds__(256) resize_general_kernel(const uint8_t* __restrict__ frames,
                                                             const StackJob* __restrict__ jobs,
                                                             uint8_t* small /* [n][16][16][16] */, uint32_t* __restrict__ done,
                                                             uint32_t* __restrict__ out_hash) {
    extern __shared__ __align__(16) uint8_t tmp[];  // ch x 16 (and the DCT cube of the stack's last frame)
    const uint32_t s = blockIdx.x >> 4, t = blockIdx.x & 15;
    const StackJob j = jobs[s];
    if (j.status != VDF_STACK_OK || j.fast) return;
    const uint8_t* img = frames + j.offset + (uint64_t)t * j.frame_stride + (uint64_t)j.top * j.pitch + j.left;
    const int32_t init_h = 1 << (j.prec_h - 1);
    for (uint32_t it = threadIdx.x; it < j.ch * 16; it += blockDim.x) {
        const uint32_t row = it >> 4, o = it & 15;
        const uint32_t s0 = j.bh[2 * o], sz = j.bh[2 * o + 1];
        const int16_t* k = j.kh + o * j.win_h;
        const uint8_t* p = img + (uint64_t)row * j.pitch + s0;
        int32_t acc = init_h;
        for (uint32_t q = 0; q < sz; ++q) acc += (int32_t)p[q] * (int32_t)k[q];
        tmp[it] = clip8(acc, j.prec_h);
    }
    __syncthreads();
    const int32_t init_v = 1 << (j.prec_v - 1);
    for (uint32_t it = threadIdx.x; it < 256; it += blockDim.x) {
        const uint32_t oy = it >> 4, ox = it & 15;
        const uint32_t s0 = j.bv[2 * oy], sz = j.bv[2 * oy + 1];
        const int16_t* k = j.kv + oy * j.win_v;
        int32_t acc = init_v;
        for (uint32_t q = 0; q < sz; ++q) acc += (int32_t)tmp[(s0 + q) * 16 + ox] * (int32_t)k[q];
        small[((uint64_t)s * 16 + t) * 256 + it] = clip8(acc, j.prec_v);
    }
    if (out_hash) finish_stack<256>(small, s, done, out_hash, reinterpret_cast<double*>(tmp), threadIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core path.  The horizontal pass is an exact integer contraction out[row][o] = sum_x k[o][x] * p[row][x]
// with u8 pixels and i16 coefficients; at ~6 multiply-adds per source byte it is far beyond what the CUDA-core
// integer pipes sustain at HBM speed (measured: IMAD/dp4a 64-92 ops/clk/SM vs 142 needed).  The i16 coefficient is
// split as k = 256*kh + kl (kh signed byte, kl unsigned byte) and the pass runs on the integer tensor path:
// A = pixels [rows x 32] u8 straight from the frame, B = coefficient bytes, 4 n-tiles = {kh, kl} x {outputs 0-7,
// 8-15}, s32 accumulators (exact: |sum| < 2^31 by fast_image_resize's own precision rule).
// One CTA per (stack, frame).  Pixels stream HBM -> shared memory in 128-byte k-chunks of kRows rows through a
// kRStages-deep cp.async ring (16-byte async copies, zero-filled outside the frame), ldmatrix feeds the A fragments,
// the coefficient fragments ride along in the same ring.  The u8 intermediate [ch][16] stays in shared memory
// (the reference rounds to u8 between the passes); the vertical pass is one thread per output pixel.
constexpr int kKch = 128;              // pixels per k-chunk (4 IMMA k-steps)
constexpr int kRowPitch = kKch + 16;   // shared-memory row pitch: ldmatrix rows land on distinct banks
constexpr int kBFragBytes = 4 * 4 * 32 * 8;  // per k-chunk: 4 k-steps x 4 n-tiles x 32 lanes x (b0,b1)

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src), "r"(src_bytes)
                 : "memory");
}
// the same through L1 (cp.async.ca): for data a thread block fetches again and again (the fused kernel's coefficient fragments)
__device__ __forceinline__ void cp_async16_ca(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void imma_u8s8(int32_t (&c)[4], const uint32_t (&a)[4], uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void imma_u8u8(int32_t (&c)[4], const uint32_t (&a)[4], uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

template <int WARPS>
struct ResizeMma {
    static constexpr int kThreads = WARPS * 32;
    static constexpr int kRows = WARPS * 32;  // rows per row block: two m16 tiles per warp
    static constexpr int kStageBytes = kRows * kRowPitch + kBFragBytes;
    static constexpr int kStages = WARPS == 4 ? 4 : 3;  // deeper ring for the small-CTA config (two CTAs per SM)
    static constexpr int kRingBytes = kStages * kStageBytes;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 8 ? 1 : 2)
    resize_mma_kernel(const uint8_t* __restrict__ frames, const StackJob* __restrict__ jobs, uint8_t* small, uint32_t* __restrict__ done,
                      uint32_t* __restrict__ out_hash) {
    using Cfg = ResizeMma<WARPS>;
    constexpr int kRowsPerPass = Cfg::kThreads / 8;        // rows covered by one 16-byte copy per thread
    constexpr int kCopies = Cfg::kRows / kRowsPerPass;     // pixel copies per thread per stage
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    __shared__ uint8_t s_mask[256];
    uint8_t* ring = smem_dyn;
    uint8_t* tmp = smem_dyn + Cfg::kRingBytes;  // ch x 16
    const uint32_t s = blockIdx.x >> 4, t = blockIdx.x & 15;
    const StackJob j = jobs[s];
    if (j.status != VDF_STACK_OK || !j.fast) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* img = frames + j.offset + (uint64_t)t * j.frame_stride + (uint64_t)j.top * j.pitch + j.x0_al;
    const uint32_t n_rb = (j.ch + Cfg::kRows - 1) / Cfg::kRows;
    const uint32_t total = n_rb * j.n_kch;

    // ---- producer: (row block, k-chunk) advance incrementally, no divisions or 64-bit multiplies per copy
    const uint32_t c16 = (tid & 7) * 16, r0 = tid >> 3;
    const uint64_t row_step = (uint64_t)kRowsPerPass * j.pitch;
    const uint8_t* g_rb = img + (uint64_t)r0 * j.pitch + c16;  // this thread's first row of the current row block
    const uint32_t sdst0 = r0 * kRowPitch + c16;
    uint32_t p_it = 0, p_kc = 0, p_row0 = r0, p_stage = 0;
    auto issue = [&]() {
        if (p_it < total) {
            uint8_t* st = ring + p_stage * Cfg::kStageBytes;
            const bool xok = j.x0_al + p_kc * kKch + c16 < j.pitch;
            const uint8_t* g = g_rb + p_kc * kKch;
            uint32_t row = p_row0;
#pragma unroll
            for (int i = 0; i < kCopies; ++i) {
                const bool ok = xok && row < j.ch;
                cp_async16(st + sdst0 + i * kRowsPerPass * kRowPitch, ok ? g : img, ok ? 16u : 0u);
                g += row_step;
                row += kRowsPerPass;
            }
            const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(j.kb) + (size_t)p_kc * kBFragBytes + tid * 16;
#pragma unroll
            for (int q = 0; q < kBFragBytes / 16 / Cfg::kThreads; ++q)
                cp_async16(st + Cfg::kRows * kRowPitch + tid * 16 + q * Cfg::kThreads * 16, bsrc + q * Cfg::kThreads * 16, 16u);
            ++p_it;
            if (++p_stage == Cfg::kStages) p_stage = 0;
            if (++p_kc == j.n_kch) {
                p_kc = 0;
                p_row0 += Cfg::kRows;
                g_rb += (uint64_t)Cfg::kRows * j.pitch;
            }
        }
        cp_async_commit();
    };

    int32_t acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0;

    for (int q = 0; q < Cfg::kStages - 1; ++q) issue();
    // tap masks of all k-chunks (<= 256 chunks) live in shared memory: no global load per iteration
    for (uint32_t q = tid; q < j.n_kch && q < 256; q += Cfg::kThreads) s_mask[q] = j.kmask[q];
    const int32_t round_h = 1 << (j.prec_h - 1);
    // ldmatrix lane -> (row, byte) inside a 16 x 32-byte A tile: lanes 0-7 rows 0-7 k 0-15, 8-15 rows 8-15 k 0-15,
    // 16-23 rows 0-7 k 16-31, 24-31 rows 8-15 k 16-31  (= registers a0..a3 of mma.m16n8k32)
    const uint32_t lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lbyte = (lane >> 4) * 16;
    const uint32_t a_off = (warp * 32 + lrow) * kRowPitch + lbyte;

    uint32_t c_kc = 0, c_rb = 0, c_stage = 0;
    for (uint32_t it = 0; it < total; ++it) {
        cp_async_wait<Cfg::kStages - 2>();
        __syncthreads();
        issue();
        const uint8_t* st = ring + c_stage * Cfg::kStageBytes;
        const uint2* sb = reinterpret_cast<const uint2*>(st + Cfg::kRows * kRowPitch);
        const uint8_t* arow = st + a_off;
        const uint32_t tapmask = s_mask[c_kc];  // the coefficient band is ~7 outputs wide
        // three straight-line bodies per k-chunk: taps only in outputs 0-7, only in 8-15, or in both
        const bool lo_oct = (tapmask & 0x55u) != 0, hi_oct = (tapmask & 0xAAu) != 0;
        if (lo_oct && hi_oct) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(a0, arow + ks * 32);
                ldmatrix_x4(a1, arow + 16 * kRowPitch + ks * 32);
                const uint2 b0 = sb[(ks * 4 + 0) * 32 + lane], b1 = sb[(ks * 4 + 1) * 32 + lane];
                const uint2 b2 = sb[(ks * 4 + 2) * 32 + lane], b3 = sb[(ks * 4 + 3) * 32 + lane];
                imma_u8s8(acc[0][0], a0, b0);
                imma_u8s8(acc[1][0], a1, b0);
                imma_u8s8(acc[0][1], a0, b1);
                imma_u8s8(acc[1][1], a1, b1);
                imma_u8u8(acc[0][2], a0, b2);
                imma_u8u8(acc[1][2], a1, b2);
                imma_u8u8(acc[0][3], a0, b3);
                imma_u8u8(acc[1][3], a1, b3);
            }
        } else if (lo_oct) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(a0, arow + ks * 32);
                ldmatrix_x4(a1, arow + 16 * kRowPitch + ks * 32);
                const uint2 b0 = sb[(ks * 4 + 0) * 32 + lane], b2 = sb[(ks * 4 + 2) * 32 + lane];
                imma_u8s8(acc[0][0], a0, b0);
                imma_u8s8(acc[1][0], a1, b0);
                imma_u8u8(acc[0][2], a0, b2);
                imma_u8u8(acc[1][2], a1, b2);
            }
        } else if (hi_oct) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t a0[4], a1[4];
                ldmatrix_x4(a0, arow + ks * 32);
                ldmatrix_x4(a1, arow + 16 * kRowPitch + ks * 32);
                const uint2 b1 = sb[(ks * 4 + 1) * 32 + lane], b3 = sb[(ks * 4 + 3) * 32 + lane];
                imma_u8s8(acc[0][1], a0, b1);
                imma_u8s8(acc[1][1], a1, b1);
                imma_u8u8(acc[0][3], a0, b3);
                imma_u8u8(acc[1][3], a1, b3);
            }
        }
        if (++c_stage == Cfg::kStages) c_stage = 0;
        if (++c_kc == j.n_kch) {  // row block finished: k = 256*kh + kl, round, shift, clamp -> u8 intermediate
            const uint32_t g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t row = c_rb * Cfg::kRows + warp * 32 + mt * 16 + g + half * 8;
#pragma unroll
                    for (int oct = 0; oct < 2; ++oct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int32_t v = acc[mt][oct][half * 2 + e] * 256 + acc[mt][2 + oct][half * 2 + e] + round_h;
                            if (row < j.ch) tmp[row * 16 + oct * 8 + q2 + e] = clip8(v, j.prec_h);
                        }
                }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][b][c] = 0;
            c_kc = 0;
            ++c_rb;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // vertical pass: the 16 x win_v coefficient table is staged in the (now idle) ring, one thread per output pixel
    int16_t* skv = reinterpret_cast<int16_t*>(ring);
    for (uint32_t q = tid; q < 16 * j.win_v; q += Cfg::kThreads) skv[q] = __ldg(j.kv + q);
    __syncthreads();
    const int32_t init_v = 1 << (j.prec_v - 1);
    for (uint32_t it = tid; it < 256; it += Cfg::kThreads) {
        const uint32_t oy = it >> 4, ox = it & 15;
        const uint32_t s0 = j.bv[2 * oy], sz = j.bv[2 * oy + 1];
        const int16_t* k = skv + oy * j.win_v;
        const uint8_t* px = tmp + s0 * 16 + ox;
        int32_t a = init_v;
#pragma unroll 8
        for (uint32_t q = 0; q < sz; ++q) a += (int32_t)px[q * 16] * (int32_t)k[q];
        small[((uint64_t)s * 16 + t) * 256 + it] = clip8(a, j.prec_v);
    }
    if (out_hash) finish_stack<Cfg::kThreads>(small, s, done, out_hash, reinterpret_cast<double*>(ring), tid);
}

// sides[n][n_fr][4] <- letterbox strip counts of frames 0, fr_step, .. of every good stack (three launches, see above)
static int letterbox_scan(vdf_ctx* ctx, cudaStream_t st, const uint8_t* d_frames, const StackDev* d_sd, uint32_t n, uint32_t n_fr, uint32_t fr_step,
                          const uint8_t* d_luts, uint32_t* d_sides) {
    if (n == 0) return VDF_OK;
    const uint32_t n_sides = n * n_fr * 4;
    VDF_ALLOC(ctx, ctx->h_lbwork.ensure(((size_t)n_sides + 1) * 4));
    uint32_t* n_work = ctx->h_lbwork.as<uint32_t>();
    uint32_t* work = n_work + 1;
    VDF_CUDA(ctx, cudaMemsetAsync(n_work, 0, 4, st));
    const unsigned persistent = (unsigned)ctx->sm_count * 4;
    if (d_luts) {
        letterbox_strip0_kernel<true><<<n_sides, 128, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, d_luts, work, n_work);
        letterbox_panels_kernel<true><<<persistent, 256, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, d_luts, work, n_work);
        letterbox_tail_kernel<true><<<persistent, 256, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, d_luts, work, n_work);
    } else {
        letterbox_strip0_kernel<false><<<n_sides, 128, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, nullptr, work, n_work);
        letterbox_panels_kernel<false><<<persistent, 256, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, nullptr, work, n_work);
        letterbox_tail_kernel<false><<<persistent, 256, 0, st>>>(d_frames, d_sd, d_sides, n_fr, fr_step, nullptr, work, n_work);
    }
    ctx->launches += 2;
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

// the letterbox union over ALL 16 frames of every stack, each pixel through the stack's table first (motion.cu)
int letterbox_all_frames(vdf_ctx* ctx, const uint8_t* d_frames, const StackDev* d_sd, uint32_t n, const uint8_t* d_luts, uint32_t* d_sides,
                         uint32_t* d_crop) {
    VDF_CUDA(ctx, cudaMemsetAsync(d_sides, 0, (size_t)n * 16 * 4 * 4, ctx->stream));  // stacks in error are skipped by the scan
    VDF_TRY(letterbox_scan(ctx, ctx->stream, d_frames, d_sd, n, 16u, 1u, d_luts, d_sides));
    crop_combine_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_sd, d_sides, n, 16u, d_crop);
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

// stand-alone DCT + pack over the stacks of a pass (context option "hash_fuse_dct" = 0): stacks the pass skipped keep what an
// earlier pass wrote, stacks in error read as zero
__global__ void __launch_bounds__(256) dct_pack_jobs_kernel(const uint8_t* __restrict__ small, const StackJob* __restrict__ jobs, uint32_t n,
                                                            uint32_t* __restrict__ out_hash) {
    __shared__ double cube[16 * 16 * 17];
    const uint32_t s = blockIdx.x;
    const int tid = threadIdx.x;
    const int32_t status = jobs[s].status;
    if (status == kJobSkip) return;
    uint32_t* out = out_hash + (uint64_t)s * 32;
    if (status != VDF_STACK_OK) {
        if (tid < 32) out[tid] = 0;
        return;
    }
    dct_pack_block<256>(small + (uint64_t)s * 4096, cube, out, tid);
}

// ================================================================================ resize jobs, built on the device
// size -> coefficient table (one axis), and (cropped width, left & 15) -> tensor-core B fragments: lookup tables in HBM that
// the host fills as it builds tables (the arithmetic is the reference's: f64, libm sin).  A zero entry = not built yet.
constexpr uint32_t kMaxDim = 16384;
struct CoefRef {
    const uint32_t* bounds;
    const int16_t* k;
    uint32_t window, precision;
    const uint4* kva;  // the same coefficients as A fragments of the vertical pass on the tensor path (fused kernel)
};
struct BFragRef {
    const uint2* kb;
    const uint8_t* kmask;
    const uint8_t* kb2;  // fused kernel's form: [256-px chunk][8 KB fragments + 16-byte header]
};

// crop -> StackJob (what the host did in round 1 between two kernels, with a synchronise in between).  *missed: a size seen for
// the first time -- the host builds its tables after the pass and the stack runs again.  max_kch / ring_bytes: limits of the
// per-frame kernels (tap masks in shared memory, the vertical table staged in the ring); the fused kernel has neither.
__device__ __forceinline__ StackJob make_job(const StackDev& d, const uint32_t* c, const CoefRef* __restrict__ coef_lut,
                                             const BFragRef* __restrict__ bfrag_lut, uint32_t ring_bytes, uint32_t max_kch, bool allow_fast,
                                             bool* missed) {
    StackJob j;
    memset(&j, 0, sizeof j);
    j.status = d.status;
    *missed = false;
    if (d.status != VDF_STACK_OK) return j;
    j.offset = d.offset, j.frame_stride = d.frame_stride, j.pitch = d.pitch;
    j.left = c[0], j.top = c[2];
    j.cw = d.width - c[0] - c[1], j.ch = d.height - c[2] - c[3];  // Crop::as_view_args, crop.rs:92-103
    const CoefRef th = coef_lut[j.cw], tv = coef_lut[j.ch];
    const uint32_t shift = j.left & 15u;
    const bool fits = (shift + j.cw + kKch - 1) / kKch <= max_kch && (uint64_t)32u * tv.window <= ring_bytes;
    const bool fast = d.aligned && fits && allow_fast;
    const BFragRef bf = fast ? bfrag_lut[(size_t)j.cw * 16 + shift] : BFragRef{nullptr, nullptr, nullptr};
    if (!th.k || !tv.k || (fast && !bf.kb)) {
        *missed = true;
        j.status = kJobSkip;
        return j;
    }
    j.bh = th.bounds, j.kh = th.k, j.win_h = th.window, j.prec_h = th.precision;
    j.bv = tv.bounds, j.kv = tv.k, j.win_v = tv.window, j.prec_v = tv.precision;
    if (fast) {
        j.x0_al = j.left & ~15u;
        j.n_kch = (shift + j.cw + kKch - 1) / kKch;
        j.kb = bf.kb, j.kmask = bf.kmask;
        j.kb2 = bf.kb2, j.kva = tv.kva, j.n_kc2 = (j.n_kch + 1) / 2;
        j.fast = 1;
    }
    return j;
}

// one thread per stack.  flags / n_nofuse (fused path, crops known up front): flags[s] <- 1 "job is built"; *n_nofuse counts the
// stacks the fused kernel will NOT hash (in error, missed, skipped, or on the general path)
__global__ void job_build_kernel(const StackDev* __restrict__ stacks, const uint32_t* __restrict__ crop, uint32_t n,
                                 const CoefRef* __restrict__ coef_lut, const BFragRef* __restrict__ bfrag_lut, uint32_t ring_bytes, uint32_t max_kch,
                                 uint32_t allow_fast, uint32_t only_missed, StackJob* __restrict__ jobs, uint32_t* __restrict__ miss,
                                 uint32_t* __restrict__ n_miss, uint32_t* __restrict__ flags, uint32_t* __restrict__ n_nofuse) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    StackDev d = stacks[s];
    StackJob j;
    if (d.status == VDF_STACK_OK && only_missed && !miss[s]) {  // hashed in the first pass
        memset(&j, 0, sizeof j);
        j.status = kJobSkip;
    } else {
        bool missed;
        j = make_job(d, crop + (size_t)s * 4, coef_lut, bfrag_lut, ring_bytes, max_kch, allow_fast != 0, &missed);
        miss[s] = missed ? 1u : 0u;
        if (missed) atomicAdd(n_miss, 1u);
    }
    jobs[s] = j;
    if (flags) {
        flags[s] = 1u;
        if (!(j.status == VDF_STACK_OK && j.fast)) atomicAdd(n_nofuse, 1u);
    }
}

// ================================================================================ the fused kernel
// ONE launch per call does letterbox -> crop -> resize job -> resize -> DCT -> threshold -> pack: sm_count persistent thread blocks of
// 512 threads (224 KB of shared memory, one block per SM), warp-specialised into five groups that never share a block-wide barrier:
//   consumers  (warps 0-7)    warp = (row group r, k-half h): rows 32 r .. 32 r + 31 of every tile, the 128 pixels of half h: the
//                             horizontal pass is the IMMA contraction of resize_mma_kernel; once per row block the pair's partial
//                             sums meet, are rounded to the u8 intermediate, and go through 768 bytes of shared memory straight into
//                             one k-step of the VERTICAL pass, also on the tensor path (A = the 16 x 32 slice of the vertical
//                             coefficients as hi / lo bytes, B = the transposed intermediate), accumulated in registers over the
//                             frame: no per-frame intermediate buffer, no serial tail per frame, shared memory independent of the
//                             frame height.  Every warp arrives on the stage's "empty" mbarrier.
//   producers  (warps 8-9)    the cp.async side of a 4-stage ring: tiles of 128 rows x 256 bytes (+ the tile's coefficient fragments,
//                             only the octets of outputs that have taps there) HBM -> shared memory, cp.async.mbarrier.arrive on the
//                             stage's "full" mbarrier.  They run ahead ACROSS frames (the ring never drains), and they may sit in a
//                             blocked copy instruction for as long as the memory system likes without holding up a tensor instruction.
//   helpers    (warps 10-13)  everything that is latency-bound: the letterbox items (strip 0 of a stack's eight sides in one round
//                             trip; walks of the sides that have a bar), crop -> resize job -> publish; then the 16^3 DCT + threshold
//                             + pack of every stack whose sixteenth frame is finished.  All of it in the shadow of the pixel stream.
//   scheduler  (warp 14)      one lane: claims the next frame, checks that its stack's job is published (else puts the frame aside
//                             and claims another), fills a slot for producers and consumers.
//   finalizer  (warp 15)      per finished frame: the four row groups' vertical sums -> round, shift, clamp -> 256 bytes of the
//                             stack's cube; the stack's sixteenth frame goes onto the helpers' DCT queue.
// Items are claimed in order from global counters and a letterbox item never waits for anything, so every wait in the kernel is
// for work that a RUNNING group has already claimed: no deadlock, whatever the number of resident blocks.  Waits are bounded
// (kFWaitNs): a kernel that would hang sets ctl[6] and ends, and the call fails loudly.
// Why: the per-frame kernels moved 5.5 TB/s (128-byte row segments, a vertical-pass tail and a launch slot per frame, a wave tail
// per launch) and the letterbox scan held the GPU alone for 7 % of a step; a token consumer with this tile shape and ring reads
// 6.7 TB/s (csrc/microbench_read.cu, profiles/r02_microbench_read.jsonl); DESIGN.md section 4 has the steps in between.
constexpr int kFRows = 128, kFCols = 256, kFPitch = kFCols + 16;
constexpr int kFStages = 4;
constexpr int kFCoefBytes = 2 * kBFragBytes + 16;               // two 128-pixel halves + header
constexpr int kFStageBytes = kFRows * kFPitch + kFCoefBytes;    // 43 024
constexpr int kFSlots = 8;
constexpr int kFSlotMasks = 72;  // frames up to 9216 pixels wide skip untapped octets; wider ones copy every fragment
constexpr int kFTmpPitch = 48;                                  // transposed intermediate: [16 outputs][32 rows + pad], per warp
constexpr int kFHelperBytes = kCubeBytes;                       // LbShared (33.5 KB) and the DCT cube (34 KB) share the helpers' arena
constexpr int kFConsThreads = 256, kFProdThreads = 64, kFHelpThreads = 128;
constexpr int kFThreads = kFConsThreads + kFProdThreads + kFHelpThreads + 64;  // + scheduler warp + finalizer warp
constexpr uint32_t kFNoFrame = 0xFFFFFFFFu;

struct FrameSlot {
    const uint8_t* img;  // first row of the crop window, first loaded column
    const uint8_t* kb2;
    const uint4* kva;
    uint32_t pitch, ch, row_bytes, n_kc, n_rb, prec_h, prec_v, out_idx /* stack * 16 + frame, or kFNoFrame: no more frames */;
    union {
        uint8_t masks[kFSlotMasks];  // tap masks of the 128-pixel k-chunks (kmask): the producers copy only the octets that have taps
        unsigned long long masks8[kFSlotMasks / 8];
    };
};
static_assert(sizeof(FrameSlot) == 56 + kFSlotMasks, "slot layout");

struct FusedSmem {  // behind the ring and the helpers' arena
    uint8_t tmpT[4][16 * kFTmpPitch];
    int32_t pairbuf[4][16 * 32];  // a consumer pair's partial sums of a row block on their way from the k-half-1 warp to the k-half-0 warp
    int32_t vred[2][256];
    FrameSlot slots[kFSlots];
    uint32_t sched_count;  // slots filled so far (scheduler -> pixel group)
    uint32_t p_ord;        // progress of the producers: 2 * (ordinal of the frame in production) + (past its middle) (-> scheduler)
    uint32_t h_bcast, h_flag;
    uint32_t abort_all;
    uint32_t fin_stop;                 // consumers are done: frames finished in total + 1
    uint32_t fin_out[2], fin_prec[2];  // finalizer: the frame whose vertical sums are in vred[b]
    uint32_t pad;
    uint64_t full[kFStages], empty[kFStages];  // mbarriers: a stage's copies have landed / its readers are done
    uint64_t fin_full[2], fin_empty[2];        // mbarriers: all eight warps' sums are in vred[b] / vred[b] is zero again
};
constexpr size_t kFusedSmemBytes = (size_t)kFStages * kFStageBytes + kFHelperBytes + sizeof(FusedSmem);

struct FusedArgs {
    const uint8_t* frames;
    const StackDev* stacks;
    StackJob* jobs;
    uint32_t n;
    uint32_t n_lb_items;  // 8 per stack (letterbox cropdetect: the kernel scans, crops and builds the jobs), 0: jobs are built already
    uint32_t* sides;      // [n][2][4]
    uint32_t* crop;       // [n][4]
    const CoefRef* coef_lut;
    const BFragRef* bfrag_lut;
    uint32_t* miss;
    uint32_t* n_miss;
    // control block, zeroed before the launch: ctl[0] frames claimed, [1] letterbox items claimed, [2] letterbox items finished,
    // [3] DCT tickets taken, [4] DCT items pushed, [5] stacks this launch will not hash, [6] a wait timed out (error),
    // [7] walk tickets taken, [12] walk items pushed, [13] strip-0 items finished; [8]-[11] statistics
    uint32_t* ctl;
    uint32_t* flags;       // [n] job of stack s is published
    uint32_t* sides_done;  // [n]
    uint32_t* done;        // [n] frames of stack s whose 16 x 16 bytes are in `small` (fused kernel) / thread blocks done (per-frame kernels)
    uint32_t* dq;          // [n] DCT queue: stack + 1
    uint32_t* wq;          // [8 n] letterbox walk queue: side item + 1
    uint8_t* small;
    uint32_t* out_hash;    // nullptr: no DCT here (context option hash_fuse_dct = 0)
    uint32_t exp;          // builds with -DVDF_TIMING_EXPERIMENTS only (VDF_FUSED_EXP): 1 = no contraction (wrong results, timing studies)
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void imma_s8u8(int32_t (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_u8u8v(int32_t (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr uint64_t kFWaitNs = 8ull * 1000 * 1000 * 1000;  // no wait in the kernel is for more than one letterbox item or one launch

// ---- helper group: letterbox items, job building, DCT items (threads 320..447)
__device__ __forceinline__ void fused_publish_job(const FusedArgs& a, uint32_t s) {  // one thread
    const StackDev d = a.stacks[s];
    uint32_t out[4] = {0, 0, 0, 0};
    if (d.status == VDF_STACK_OK) {  // crop_combine_kernel's rule for frames 0 and 8
        const int W = (int)d.width, H = (int)d.height;
        for (uint32_t fr = 0; fr < 2; ++fr) {
            uint32_t c[4];
            for (int k = 0; k < 4; ++k) c[k] = __ldcg(&a.sides[((size_t)s * 2 + fr) * 4 + k]);
            if (!(W - (int)c[0] - (int)c[1] >= 1 && H - (int)c[2] - (int)c[3] >= 1)) c[0] = c[1] = c[2] = c[3] = 0;
            for (int k = 0; k < 4; ++k) out[k] = fr == 0 ? c[k] : min(out[k], c[k]);
        }
    }
    for (int k = 0; k < 4; ++k) a.crop[(size_t)s * 4 + k] = out[k];
    bool missed;
    const StackJob j = make_job(d, out, a.coef_lut, a.bfrag_lut, 0xFFFFFFFFu, 0xFFFFFFFFu, true, &missed);
    a.miss[s] = missed ? 1u : 0u;
    if (missed) atomicAdd(a.n_miss, 1u);
    a.jobs[s] = j;
    if (!(j.status == VDF_STACK_OK && j.fast)) atomicAdd(&a.ctl[5], 1u);
    __threadfence();
    atomicExch(&a.flags[s], 1u);
}

__device__ __forceinline__ void fused_helper(const FusedArgs& a, uint8_t* arena, FusedSmem& fs, uint64_t t_start) {
    constexpr int NT = 128, BAR = 2;
    LbShared& sh = *reinterpret_cast<LbShared*>(arena);
    const int gtid = threadIdx.x - kFConsThreads - kFProdThreads;
    auto finish_side = [&](uint32_t b, uint32_t count) {  // one thread
        const uint32_t s = b >> 3;
        a.sides[b] = count;
        __threadfence();
        if (atomicAdd(&a.sides_done[s], 1u) == 7u) {
            __threadfence();
            fused_publish_job(a, s);
        }
        __threadfence();
        atomicAdd(&a.ctl[2], 1u);
    };
    // ---- letterbox items.  Two kinds.  "Strip 0 of all eight sides of stack s" (frames 0 and 8 x left, right, top, bottom), claimed in
    // order: every load of the item is in flight at once -- one round trip to memory -- and a stack with no bar at all (most) is
    // published right there, so nearly all of the pixel work is available within the first microseconds of the launch.  A side whose
    // strip 0 is letterbox is queued as a "walk side b inwards, panel after panel" item; groups take walks once the stack items are
    // all claimed, and the last walk of a stack publishes it.  Neither kind ever waits for anything.
    const uint32_t n_stack_items = a.n_lb_items / 8u;
    for (; a.n_lb_items;) {
        group_sync<BAR, NT>();
        if (gtid == 0) {
            uint32_t item = 0;  // 0: nothing left; (s + 1): strip 0 of stack s; 0x80000000 | (b + 1): walk side b
            for (;;) {
                if (ld_volatile_u32(&a.ctl[1]) < n_stack_items) {
                    const uint32_t s = atomicAdd(&a.ctl[1], 1u);
                    if (s < n_stack_items) {
                        item = s + 1u;
                        break;
                    }
                }
                const uint32_t head = ld_volatile_u32(&a.ctl[7]);
                if (head < ld_acquire_u32(&a.ctl[12])) {
                    if (atomicCAS(&a.ctl[7], head, head + 1u) != head) continue;
                    uint32_t v;
                    while (!(v = ld_volatile_u32(&a.wq[head]))) {}  // the pusher is between its two instructions
                    item = 0x80000000u | v;
                    break;
                }
                if (ld_acquire_u32(&a.ctl[13]) >= n_stack_items && ld_volatile_u32(&a.ctl[7]) >= ld_volatile_u32(&a.ctl[12])) break;  // all found, all taken
                __nanosleep(200);
            }
            fs.h_bcast = item;
        }
        group_sync<BAR, NT>();
        const uint32_t item = fs.h_bcast;
        if (!item) break;
        __threadfence();
        if (!(item & 0x80000000u)) {
            const uint32_t s = item - 1u;
            const StackDev sd = a.stacks[s];
            uint32_t bars = 0;  // bit (fr * 4 + side): strip 0 is letterbox
            if (sd.status == VDF_STACK_OK) {
                const uint32_t W = sd.width, H = sd.height, P = sd.pitch;
                const uint8_t* f0 = a.frames + sd.offset;
                uint32_t* hist8 = sh.hist;  // 8 x 256 counters
                if (H <= NT * 9 && W <= NT * 16 && ((reinterpret_cast<uintptr_t>(f0) | P | sd.frame_stride) & 3) == 0) {
                    for (int q = gtid; q < 8 * 256; q += NT) hist8[q] = 0;
                    // columns: 9 pixels per thread and side; rows: 4 words per thread and side -- 52 loads in flight per thread
                    uint32_t vc[4][9], vr[4][4];
                    const uint32_t W4 = W >> 2;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint8_t* img = f0 + (uint64_t)((k >> 1) * 8) * sd.frame_stride;
                        const uint8_t* col = img + ((k & 1) ? W - 1 : 0u);
#pragma unroll
                        for (int u = 0; u < 9; ++u) vc[k][u] = gtid + NT * u < (int)H ? (uint32_t)__ldg(col + (uint64_t)(gtid + NT * u) * P) : 0x100u;
                        const uint32_t* row4 = reinterpret_cast<const uint32_t*>(img + (uint64_t)((k & 1) ? H - 1 : 0u) * P);
#pragma unroll
                        for (int u = 0; u < 4; ++u) vr[k][u] = gtid + NT * u < (int)W4 ? __ldg(row4 + gtid + NT * u) : 0u;
                    }
                    group_sync<BAR, NT>();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t* hc = hist8 + ((k >> 1) * 4 + (k & 1)) * 256;      // side 0 / 1 of frame k >> 1
                        uint32_t* hr = hist8 + ((k >> 1) * 4 + 2 + (k & 1)) * 256;  // side 2 / 3
#pragma unroll
                        for (int u = 0; u < 9; ++u)
                            if (vc[k][u] < 0x100u) atomicAdd(&hc[vc[k][u]], 1u);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (gtid + NT * u < (int)W4) {
                                atomicAdd(&hr[vr[k][u] & 255u], 1u);
                                atomicAdd(&hr[(vr[k][u] >> 8) & 255u], 1u);
                                atomicAdd(&hr[(vr[k][u] >> 16) & 255u], 1u);
                                atomicAdd(&hr[vr[k][u] >> 24], 1u);
                            }
                        if (gtid < (int)(W & 3u)) {  // the <= 3 pixels after the last whole word of the row
                            const uint8_t* img = f0 + (uint64_t)((k >> 1) * 8) * sd.frame_stride;
                            atomicAdd(&hr[__ldg(img + (uint64_t)((k & 1) ? H - 1 : 0u) * P + (W4 << 2) + gtid)], 1u);
                        }
                    }
                    if (gtid == 0) fs.h_flag = 0;
                    group_sync<BAR, NT>();
                    const int warp = gtid >> 5, lane = gtid & 31;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k8 = warp * 2 + e;  // fr * 4 + side
                        const bool ok = strip_is_letterbox(hist8 + k8 * 256, (k8 & 2) ? W : H, lane);
                        if (ok && lane == 0) atomicOr(&fs.h_flag, 1u << k8);
                    }
                    group_sync<BAR, NT>();
                    bars = fs.h_flag;
                } else {  // large or unaligned frames: side after side
                    for (uint32_t k8 = 0; k8 < 8; ++k8) {
                        const uint8_t* img = f0 + (uint64_t)((k8 >> 2) * 8) * sd.frame_stride;
                        if (lb_strip0<false, NT, BAR>(sh.hist, &fs.h_flag, nullptr, img, W, H, P, k8 & 3)) bars |= 1u << k8;
                        group_sync<BAR, NT>();
                    }
                }
            }
            if (gtid == 0) {
                if (!bars) {
                    for (uint32_t k8 = 0; k8 < 8; ++k8) a.sides[s * 8 + k8] = 0u;
                    __threadfence();
                    fused_publish_job(a, s);
                    __threadfence();
                    atomicAdd(&a.ctl[2], 8u);
                } else {
                    for (uint32_t k8 = 0; k8 < 8; ++k8)
                        if (!(bars >> k8 & 1u)) a.sides[s * 8 + k8] = 0u;
                    atomicExch(&a.sides_done[s], 8u - (uint32_t)__popc(bars));  // the walks count the rest
                    __threadfence();
                    atomicAdd(&a.ctl[2], 8u - (uint32_t)__popc(bars));
                    const uint32_t idx = atomicAdd(&a.ctl[12], (uint32_t)__popc(bars));
                    uint32_t q = 0;
                    for (uint32_t k8 = 0; k8 < 8; ++k8)
                        if (bars >> k8 & 1u) atomicExch(&a.wq[idx + q++], s * 8 + k8 + 1u);
                }
                __threadfence();
                atomicAdd(&a.ctl[13], 1u);
            }
        } else {
            const uint32_t b = (item & 0x7FFFFFFFu) - 1u, side = b & 3, fr = (b >> 2) & 1, s = b >> 3;
            const StackDev sd = a.stacks[s];
            const uint8_t* img = a.frames + sd.offset + (uint64_t)(fr * 8) * sd.frame_stride;
            const uint32_t limit = side < 2 ? sd.width : sd.height;
            uint32_t first = kLbNone;
            for (uint32_t base = 0; base < limit && first == kLbNone; base += kColPanel) {
                lb_panel_flags<false, NT, BAR>(sh, img, sd.width, sd.height, sd.pitch, side, base);
                if (gtid == 0) fs.h_bcast = lb_first_stop(sh, base, limit);
                group_sync<BAR, NT>();
                first = fs.h_bcast;
                group_sync<BAR, NT>();
            }
            if (gtid == 0) finish_side(b, min(first, limit));
        }
    }
    if (gtid == 0 && a.n_lb_items) atomicMax(&a.ctl[10], (uint32_t)((globaltimer_ns() - t_start) >> 10));  // statistics: end of the scan
    if (!a.out_hash) return;
    // ---- DCT items, by ticket
    double* cube = reinterpret_cast<double*>(arena);
    for (;;) {
        group_sync<BAR, NT>();
        if (gtid == 0) {
            const uint32_t idx = atomicAdd(&a.ctl[3], 1u);
            uint32_t v = 0;
            const uint64_t t0 = globaltimer_ns();
            for (uint32_t spin = 0;; ++spin) {
                if (idx < a.n) v = ld_volatile_u32(&a.dq[idx]);
                if (v) break;
                if (ld_volatile_u32(&a.ctl[6])) break;
                if (ld_acquire_u32(&a.ctl[2]) >= a.n_lb_items) {  // every job is built: the number of stacks to hash is final
                    const uint32_t skipped = ld_volatile_u32(&a.ctl[5]);
                    if (idx >= a.n - min(skipped, a.n)) break;
                }
                __nanosleep(400);
                if ((spin & 1023u) == 1023u && globaltimer_ns() - t0 > 8 * kFWaitNs) {
                    atomicExch(&a.ctl[6], 3u);
                    break;
                }
            }
            fs.h_bcast = v;
        }
        group_sync<BAR, NT>();
        const uint32_t v = fs.h_bcast;
        if (!v) break;
        __threadfence();
        dct_pack_block<NT, BAR>(a.small + (uint64_t)(v - 1) * 4096, cube, a.out_hash + (uint64_t)(v - 1) * 32, gtid);
    }
}

// ---- scheduler: one lane.  Frames are claimed in order, but a frame whose stack's job is not published yet (its letterbox items
// are still walking a bar) is put aside -- up to six per block -- and the next frame is claimed instead: nobody idles behind a slow scan.
__device__ __forceinline__ void fused_scheduler(const FusedArgs& a, FusedSmem& fs) {
    volatile uint32_t* v_pord = &fs.p_ord;
    volatile uint32_t* v_abort = &fs.abort_all;
    uint32_t cur_s = 0xFFFFFFFFu;
    StackJob job;
    memset(&job, 0, sizeof job);
    job.status = kJobSkip;
    const uint32_t total = a.n * 16u;
    constexpr int kAside = 6;
    uint32_t aside[kAside];
    for (int q = 0; q < kAside; ++q) aside[q] = kFNoFrame;
    bool exhausted = false;
    uint32_t mask_s = 0xFFFFFFFFu;
    unsigned long long cur_masks[kFSlotMasks / 8];
    for (uint32_t j = 0;; ++j) {
        // A claimed frame is a frame no other block can take: the next one is claimed only when the producers are past the middle of
        // the current one (p_prog = 2 * ordinal + past-the-middle), early enough to hide this lane's few dependent round trips to memory
        // and late enough that the blocks finish within half a frame of each other at the end of the launch.
        const uint32_t gate = j >= 1 ? 2u * j - 1u : 0u;
        while (gate > *v_pord) {
            if (*v_abort) return;
            __nanosleep(2000);  // half a frame is ~20 us away
        }
        FrameSlot sl;
        memset(&sl, 0, sizeof sl);
        sl.out_idx = kFNoFrame;
        uint64_t t0 = 0;
        for (uint32_t spin = 0;;) {
            uint32_t f = kFNoFrame;
            int free_q = -1, n_aside = 0;
            uint32_t fl[kAside];  // all flags of the frames put aside in one round trip (plain loads; the fence below orders the job read)
#pragma unroll
            for (int q = 0; q < kAside; ++q) fl[q] = aside[q] == kFNoFrame ? 0u : ((aside[q] >> 4) == cur_s ? 1u : ld_volatile_u32(&a.flags[aside[q] >> 4]));
#pragma unroll
            for (int q = 0; q < kAside; ++q) {  // a frame put aside whose job has arrived in the meantime
                if (aside[q] == kFNoFrame) {
                    free_q = q;
                    continue;
                }
                if (f == kFNoFrame && fl[q]) f = aside[q], aside[q] = kFNoFrame, free_q = q;
                else ++n_aside;
            }
            if (f != kFNoFrame) __threadfence();
            const bool room = free_q >= 0;
            if (f == kFNoFrame && !exhausted && room) {
                f = atomicAdd(&a.ctl[0], 1u);
                if (f >= total) {
                    exhausted = true;
                    continue;
                }
                if (!((f >> 4) == cur_s || ld_acquire_u32(&a.flags[f >> 4]))) {
#pragma unroll
                    for (int q = 0; q < kAside; ++q)
                        if (q == free_q) aside[q] = f;
                    continue;
                }
            }
            if (f == kFNoFrame) {
                if (exhausted && n_aside == 0) break;  // nothing left anywhere
                if (spin == 0) t0 = globaltimer_ns();
                __nanosleep(200);
                if (ld_volatile_u32(&a.ctl[6]) || ((++spin & 1023u) == 0u && globaltimer_ns() - t0 > kFWaitNs)) {
                    atomicExch(&a.ctl[6], 1u);
                    break;
                }
                continue;
            }
            if (spin) atomicAdd(&a.ctl[8], (uint32_t)((globaltimer_ns() - t0) >> 10)), spin = 0;  // statistics: ~us with nothing ready
            const uint32_t s = f >> 4;
            if (s != cur_s) {
                const volatile uint64_t* src = reinterpret_cast<const volatile uint64_t*>(&a.jobs[s]);  // written by another SM
                uint64_t* dst = reinterpret_cast<uint64_t*>(&job);
#pragma unroll
                for (int q = 0; q < (int)(sizeof(StackJob) / 8); ++q) dst[q] = src[q];
                cur_s = s;
            }
            if (job.status == VDF_STACK_OK && job.fast) {
                const uint32_t t = f & 15u;
                sl.img = a.frames + job.offset + (uint64_t)t * job.frame_stride + (uint64_t)job.top * job.pitch + job.x0_al;
                sl.kb2 = job.kb2, sl.kva = job.kva;
                sl.pitch = job.pitch, sl.ch = job.ch, sl.row_bytes = job.pitch - job.x0_al, sl.n_kc = job.n_kc2;
                sl.n_rb = (job.ch + kFRows - 1) / kFRows, sl.prec_h = job.prec_h, sl.prec_v = job.prec_v;
                sl.out_idx = f;
                if (s != mask_s) {  // the table's masks (static data, 4096-byte aligned, followed by the kb2 blob): nine words, all in flight
                    const bool have = job.n_kch <= (uint32_t)kFSlotMasks;
                    const unsigned long long* km = reinterpret_cast<const unsigned long long*>(job.kmask);
#pragma unroll
                    for (int w = 0; w < kFSlotMasks / 8; ++w) cur_masks[w] = __ldg(km + w);
#pragma unroll
                    for (int w = 0; w < kFSlotMasks / 8; ++w) {
                        const uint32_t left = job.n_kch > (uint32_t)(8 * w) ? job.n_kch - 8 * w : 0u;  // bytes of this word that are masks
                        const unsigned long long keep = left >= 8 ? ~0ull : ((1ull << (8 * left)) - 1ull);
                        cur_masks[w] = have ? (cur_masks[w] & keep) : ~0ull;
                    }
                    mask_s = s;
                }
#pragma unroll
                for (int w = 0; w < kFSlotMasks / 8; ++w) sl.masks8[w] = cur_masks[w];
                break;
            }
        }
        fs.slots[j % kFSlots] = sl;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t*>(&fs.sched_count) = j + 1;
        if (sl.out_idx == kFNoFrame) return;
    }
}

// mbarrier helpers (shared::cta)
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
// this thread's arrival happens when all of its earlier cp.async copies have landed
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
            (uint32_t)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}

// ---- finalizer (one warp): per finished frame, the sum of the eight consumer warps' vertical partial sums -> round, shift, clamp ->
// 256 bytes of the stack's 16 x 16 x 16 cube in global memory; the stack's sixteenth frame goes onto the helpers' DCT queue.
__device__ __forceinline__ void fused_finalizer(const FusedArgs& a, FusedSmem& fs) {
    const int lane = threadIdx.x & 31;
    volatile uint32_t* v_stop = &fs.fin_stop;
    for (uint32_t ord = 0;; ++ord) {
        const uint32_t b = ord & 1u, parity = (ord >> 1) & 1u;
        // wait for frame `ord` or for the end
        for (;;) {
            uint32_t ready;  // suspended in hardware for up to ~4 us at a time: this warp costs no issue slots while it waits
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ready)
                         : "r"((uint32_t)__cvta_generic_to_shared(&fs.fin_full[b])), "r"(parity), "r"(4000u)
                         : "memory");
            if (ready) break;
            const uint32_t stop = *v_stop;
            if (stop && ord + 1u >= stop) return;
            if (*reinterpret_cast<volatile uint32_t*>(&fs.abort_all)) return;
        }
        const uint32_t out = fs.fin_out[b], prec = fs.fin_prec[b];
        int32_t* vr = fs.vred[b];
        const int32_t init_v = 1 << (prec - 1);
        uint32_t w[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int4 v = *reinterpret_cast<const int4*>(vr + lane * 8 + h * 4);
            *reinterpret_cast<int4*>(vr + lane * 8 + h * 4) = make_int4(0, 0, 0, 0);
            w[h] = (uint32_t)clip8(v.x + init_v, prec) | ((uint32_t)clip8(v.y + init_v, prec) << 8) | ((uint32_t)clip8(v.z + init_v, prec) << 16) |
                   ((uint32_t)clip8(v.w + init_v, prec) << 24);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&fs.fin_empty[b]);
        *reinterpret_cast<uint2*>(a.small + (uint64_t)out * 256 + lane * 8) = make_uint2(w[0], w[1]);
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            const uint32_t s = out >> 4;
            if (atomicAdd(&a.done[s], 1u) == 15u && a.out_hash) {
                __threadfence();
                const uint32_t idx = atomicAdd(&a.ctl[4], 1u);
                atomicExch(&a.dq[idx], s + 1u);
            }
        }
    }
}

// ---- producer group (threads 256..319): the cp.async side of the ring.  It may sit in a blocked copy instruction for as long as the
// memory system likes (that is what back-pressure looks like) without holding up a single tensor instruction of the consumers.
__device__ __forceinline__ void fused_producer(const FusedArgs& a, uint8_t* ring, FusedSmem& fs) {
    constexpr int NT = kFProdThreads;
    const int tid = threadIdx.x - kFConsThreads;
    volatile uint32_t* v_sched = &fs.sched_count;
    // 16 threads x 16 bytes cover a 256-byte row segment, 4 rows per pass, 32 passes per tile
    const uint32_t c16 = (tid & 15) * 16, r0 = tid >> 4;
    uint32_t stage = 0, empty_parity = 1;  // a fresh mbarrier passes a wait on the phase before its first
    for (uint32_t ord = 0;; ++ord) {
        const uint64_t t0 = globaltimer_ns();
        uint32_t spin = 0;
        for (; *v_sched <= ord; ++spin) {
            __nanosleep(40);
            if ((spin & 4095u) == 4095u && globaltimer_ns() - t0 > 2 * kFWaitNs) {
                atomicExch(&a.ctl[6], 2u);
                fs.abort_all = 1;
                return;
            }
        }
        if (spin && tid == 0) atomicAdd(&a.ctl[ord == 0 ? 14 : 9], (uint32_t)((globaltimer_ns() - t0) >> 10));  // statistics: ~us the ring waited for its first / a later slot
        __threadfence_block();
        const FrameSlot& sl = fs.slots[ord % kFSlots];
        if (sl.out_idx == kFNoFrame) return;
        const uint8_t* img = sl.img;
        const uint8_t* kb2 = sl.kb2;
        const uint8_t* pm = sl.masks;
        // this thread's eight coefficient copies of a tile are all of one n-tile = one octet of outputs: ((tid >> 4) & 1); copies q = 0..3
        // belong to the tile's first 128-pixel half, q = 4..7 to the second
        const uint32_t oct_bits = 0x55u << ((tid >> 4) & 1);
        const uint32_t n_kc = sl.n_kc, n_rb = sl.n_rb, ch = sl.ch, row_bytes = sl.row_bytes;
        const uint64_t row_step = (uint64_t)4 * sl.pitch;
        if (tid == 0) *reinterpret_cast<volatile uint32_t*>(&fs.p_ord) = 2u * ord + (n_rb < 2u ? 1u : 0u);
        const uint8_t* g_rb = img + (uint64_t)r0 * sl.pitch + c16;
        for (uint32_t rb = 0; rb < n_rb; ++rb, g_rb += 32 * row_step) {
            if (tid == 0 && rb == n_rb / 2 && rb) *reinterpret_cast<volatile uint32_t*>(&fs.p_ord) = 2u * ord + 1u;
            for (uint32_t kci = 0; kci < n_kc; ++kci) {
                // k-chunks left to right on even row blocks, right to left on odd ones: the coefficient fragments a row block ends with are
                // the ones the next one starts with, and they come through L1 (what shared memory leaves of it holds a few chunks)
                const uint32_t kc = (rb & 1u) ? n_kc - 1u - kci : kci;
                mbar_wait(&fs.empty[stage], empty_parity);
                uint8_t* st = ring + (size_t)stage * kFStageBytes;
                const bool xok = kc * kFCols + c16 < row_bytes;
                const uint8_t* g = g_rb + kc * kFCols;
                uint32_t row = rb * kFRows + r0;
                uint8_t* d = st + r0 * kFPitch + c16;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const bool ok = xok && row < ch;
                    cp_async16(d + i * 4 * kFPitch, ok ? g : img, ok ? 16u : 0u);
                    g += row_step;
                    row += 4;
                }
                const uint8_t* csrc = kb2 + (size_t)kc * kFCoefBytes + tid * 16;
                uint8_t* cdst = st + kFRows * kFPitch + tid * 16;
                const bool wide = 2 * kc + 1 >= (uint32_t)kFSlotMasks;
                const bool need0 = wide || (pm[2 * kc] & oct_bits), need1 = wide || (pm[2 * kc + 1] & oct_bits);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < 4 ? need0 : need1) cp_async16_ca(cdst + q * NT * 16, csrc + q * NT * 16);
                if (tid == 0) cp_async16_ca(cdst + 8 * NT * 16, csrc + 8 * NT * 16);  // the header
                mbar_arrive_on_cp_async(&fs.full[stage]);
                if (++stage == kFStages) stage = 0, empty_parity ^= 1u;
            }
        }
    }
}

__global__ void __launch_bounds__(kFThreads, 1) hash_fused_kernel(const FusedArgs a) {
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    uint8_t* ring = smem_dyn;
    uint8_t* arena = smem_dyn + (size_t)kFStages * kFStageBytes;
    FusedSmem& fs = *reinterpret_cast<FusedSmem*>(arena + kFHelperBytes);
    const int tid = threadIdx.x;
    const uint64_t t_start = globaltimer_ns();
    for (int q = tid; q < 512; q += kFThreads) (&fs.vred[0][0])[q] = 0;
    if (tid == 0) {
        fs.sched_count = 0, fs.p_ord = 0, fs.abort_all = 0;
        for (int s = 0; s < kFStages; ++s) mbar_init(&fs.full[s], kFProdThreads), mbar_init(&fs.empty[s], kFConsThreads / 32);
        for (int s = 0; s < 2; ++s) mbar_init(&fs.fin_full[s], kFConsThreads / 64), mbar_init(&fs.fin_empty[s], 1);
        fs.fin_stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= kFConsThreads + kFProdThreads + kFHelpThreads + 32) {
        fused_finalizer(a, fs);
        return;
    }
    if (tid >= kFConsThreads + kFProdThreads + kFHelpThreads) {
        if (tid == kFConsThreads + kFProdThreads + kFHelpThreads) fused_scheduler(a, fs);
        return;
    }
    if (tid >= kFConsThreads + kFProdThreads) {
        fused_helper(a, arena, fs, t_start);
        return;
    }
    if (tid >= kFConsThreads) {
        fused_producer(a, ring, fs);
        return;
    }
    // ------------------------------------------------------------------------------------------------ consumer group
    const int lane = tid & 31, warp = tid >> 5;
    volatile uint32_t* v_sched = &fs.sched_count;
    // Warp w = (row group r = w >> 1, k-half h = w & 1): rows 32 r .. 32 r + 31 of every tile (two m16 tiles), the 128 pixels of half h,
    // all sixteen outputs x {high, low} coefficient bytes.  A coefficient fragment is read from shared memory by four warps (not
    // eight: the consumers' shared-memory traffic is what limits the kernel at the clock the power cap leaves in long runs); the
    // two warps of a pair add their partial sums once per row block, before the rounding.
    int32_t acc[2][4][4], vacc[2][2][4];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int z = 0; z < 4; ++z) acc[x][y][z] = 0;
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int z = 0; z < 4; ++z) vacc[x][y][z] = 0;
    const uint32_t rgrp = warp >> 1, khalf = warp & 1;
    const uint32_t lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lbyte = (lane >> 4) * 16;
    const uint32_t a_off = (rgrp * 32 + lrow) * kFPitch + lbyte + khalf * kKch;
    const uint32_t g = lane >> 2, q4 = lane & 3, q2 = q4 * 2;
    uint8_t* tw = fs.tmpT[rgrp];
    int32_t* pb = fs.pairbuf[rgrp] + lane;
    // pair barriers (64 threads): A = "the partial sums are in pairbuf", B = "pairbuf has been read"
    auto pair_arrive = [&](uint32_t id) {
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
    };
    auto pair_sync = [&](uint32_t id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); };
    const uint32_t bar_a = 3u + 2u * rgrp, bar_b = 4u + 2u * rgrp;
    bool pb_busy = false;  // (k-half 1) the partner has not yet confirmed reading the previous row block's sums

    uint32_t c_ord = 0, c_kc = 0, c_rb = 0, c_stage = 0, full_parity = 0;
    uint32_t c_nkc = 0, c_nrb = 0, c_prec_h = 0, c_prec_v = 0, c_out = kFNoFrame;
    const uint4* c_kva = nullptr;
    auto c_take_slot = [&]() {
        const uint64_t t0 = globaltimer_ns();
        for (uint32_t spin = 0; *v_sched <= c_ord; ++spin) {  // only the first frame can make the consumers wait: the producers go first
            __nanosleep(40);
            if ((spin & 4095u) == 4095u && globaltimer_ns() - t0 > 2 * kFWaitNs) {
                c_out = kFNoFrame;
                fs.abort_all = 1;
                return;
            }
        }
        __threadfence_block();
        const FrameSlot& sl = fs.slots[c_ord % kFSlots];
        c_out = sl.out_idx;
        c_nkc = sl.n_kc, c_nrb = sl.n_rb, c_prec_h = sl.prec_h, c_prec_v = sl.prec_v, c_kva = sl.kva;
        c_kc = 0, c_rb = 0;
    };
    c_take_slot();

    while (c_out != kFNoFrame) {
        mbar_wait(&fs.full[c_stage], full_parity);
        const uint8_t* st = ring + (size_t)c_stage * kFStageBytes;
        const uint8_t* coef = st + kFRows * kFPitch;
        const uint32_t masks = *reinterpret_cast<const uint32_t*>(coef + 2 * kBFragBytes);
        const bool last_kc = c_kc + 1 == c_nkc;
        uint4 va_h = make_uint4(0, 0, 0, 0), va_l = make_uint4(0, 0, 0, 0);
        if (last_kc && khalf == 0) {  // the vertical k-step's A fragments: in flight under the row block's last contraction
            const uint4* p = c_kva + ((size_t)(c_rb * 4 + rgrp) * 2) * 32 + lane;
            va_h = __ldg(p), va_l = __ldg(p + 32);
        }
        {
            const uint8_t* arow = st + a_off;
            const uint2* sb = reinterpret_cast<const uint2*>(coef + khalf * kBFragBytes);
#ifdef VDF_TIMING_EXPERIMENTS  // make EXTRA=-DVDF_TIMING_EXPERIMENTS + VDF_FUSED_EXP=1: the launch without its contraction (WRONG hashes; timing studies)
            const uint32_t tapmask = (a.exp & 1u) ? 0u : (masks >> (8 * khalf)) & 0xFFu;
#else
            const uint32_t tapmask = (masks >> (8 * khalf)) & 0xFFu;
#endif
            const bool lo_oct = (tapmask & 0x55u) != 0, hi_oct = (tapmask & 0xAAu) != 0;
            if (lo_oct && hi_oct) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, arow + ks * 32);
                    ldmatrix_x4(a1, arow + 16 * kFPitch + ks * 32);
                    const uint2 b0 = sb[(ks * 4 + 0) * 32 + lane], b1 = sb[(ks * 4 + 1) * 32 + lane];
                    const uint2 b2 = sb[(ks * 4 + 2) * 32 + lane], b3 = sb[(ks * 4 + 3) * 32 + lane];
                    imma_u8s8(acc[0][0], a0, b0);
                    imma_u8s8(acc[1][0], a1, b0);
                    imma_u8s8(acc[0][1], a0, b1);
                    imma_u8s8(acc[1][1], a1, b1);
                    imma_u8u8(acc[0][2], a0, b2);
                    imma_u8u8(acc[1][2], a1, b2);
                    imma_u8u8(acc[0][3], a0, b3);
                    imma_u8u8(acc[1][3], a1, b3);
                }
            } else if (lo_oct) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, arow + ks * 32);
                    ldmatrix_x4(a1, arow + 16 * kFPitch + ks * 32);
                    const uint2 b0 = sb[(ks * 4 + 0) * 32 + lane], b2 = sb[(ks * 4 + 2) * 32 + lane];
                    imma_u8s8(acc[0][0], a0, b0);
                    imma_u8s8(acc[1][0], a1, b0);
                    imma_u8u8(acc[0][2], a0, b2);
                    imma_u8u8(acc[1][2], a1, b2);
                }
            } else if (hi_oct) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, arow + ks * 32);
                    ldmatrix_x4(a1, arow + 16 * kFPitch + ks * 32);
                    const uint2 b1 = sb[(ks * 4 + 1) * 32 + lane], b3 = sb[(ks * 4 + 3) * 32 + lane];
                    imma_u8s8(acc[0][1], a0, b1);
                    imma_u8s8(acc[1][1], a1, b1);
                    imma_u8u8(acc[0][3], a0, b3);
                    imma_u8u8(acc[1][3], a1, b3);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&fs.empty[c_stage]);  // this warp has read everything it needs from the stage
        if (++c_stage == kFStages) c_stage = 0, full_parity ^= 1u;
        if (last_kc) {
            // row block finished: this warp's partial sums k = 256 kh + kl over its k-half, 16 per lane
            int32_t v[2][2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int oct = 0; oct < 2; ++oct)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[mt][oct][j] = acc[mt][oct][j] * 256 + acc[mt][2 + oct][j];
                        acc[mt][oct][j] = 0, acc[mt][2 + oct][j] = 0;
                    }
            if (khalf == 1) {
                if (pb_busy) pair_sync(bar_b);  // the partner's arrival after reading the previous row block's sums
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int oct = 0; oct < 2; ++oct)
#pragma unroll
                        for (int j = 0; j < 4; ++j) pb[((mt * 2 + oct) * 4 + j) * 32] = v[mt][oct][j];
                pair_arrive(bar_a);
                pb_busy = true;
            } else {
                pair_sync(bar_a);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int oct = 0; oct < 2; ++oct)
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[mt][oct][j] += pb[((mt * 2 + oct) * 4 + j) * 32];
                pair_arrive(bar_b);
                // round, shift, clamp -> the u8 intermediate of the pair's 32 rows, transposed
                const int32_t round_h = 1 << (c_prec_h - 1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t row = mt * 16 + half * 8 + g;
#pragma unroll
                        for (int oct = 0; oct < 2; ++oct)
#pragma unroll
                            for (int e = 0; e < 2; ++e) tw[(oct * 8 + q2 + e) * kFTmpPitch + row] = clip8(v[mt][oct][half * 2 + e] + round_h, c_prec_h);
                    }
                __syncwarp();
                // one k-step (these 32 rows) of the vertical pass: out[oy][ox] += sum_y kv[oy][y] * tmp[y][ox]
                const uint32_t ah[4] = {va_h.x, va_h.y, va_h.z, va_h.w}, al[4] = {va_l.x, va_l.y, va_l.z, va_l.w};
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const uint8_t* bp = tw + (nt * 8 + g) * kFTmpPitch + 4 * q4;
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(bp), b1 = *reinterpret_cast<const uint32_t*>(bp + 16);
                    imma_s8u8(vacc[nt][0], ah, b0, b1);
                    imma_u8u8v(vacc[nt][1], al, b0, b1);
                }
                __syncwarp();
            }
            c_kc = 0;
            if (++c_rb == c_nrb) {  // frame finished: the four row groups' vertical sums meet in shared memory; the finalizer warp clamps,
                if (khalf == 0) {    // stores and counts -- no barrier and no memory fence on the consumers' path
                    int32_t* vr = fs.vred[c_ord & 1];
                    mbar_wait(&fs.fin_empty[c_ord & 1], ((c_ord >> 1) & 1u) ^ 1u);  // the finalizer is done with frame c_ord - 2 (passes at once, normally)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int half = 0; half < 2; ++half)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int32_t sum = vacc[nt][0][half * 2 + e] * 256 + vacc[nt][1][half * 2 + e];
                                atomicAdd(&vr[(g + half * 8) * 16 + nt * 8 + q2 + e], sum);
                                vacc[nt][0][half * 2 + e] = 0, vacc[nt][1][half * 2 + e] = 0;
                            }
                    if (tid == 0) fs.fin_out[c_ord & 1] = c_out, fs.fin_prec[c_ord & 1] = c_prec_v;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&fs.fin_full[c_ord & 1]);
                }
                ++c_ord;
                c_take_slot();
            }
        } else {
            ++c_kc;
        }
    }
    if (khalf == 1 && pb_busy) pair_sync(bar_b);  // pair barriers end balanced
    if (tid == 0) {
        atomicAdd(&a.ctl[11], (uint32_t)((globaltimer_ns() - t_start) >> 10));  // statistics
        *reinterpret_cast<volatile uint32_t*>(&fs.fin_stop) = c_ord + 1u;      // every thread has the same c_ord: frames this block finished
    }
}

// ================================================================================ host side
// fast_image_resize 5.1 (third-party, un-vendored): Lanczos3, adaptive kernel size, coefficients
// normalised in f64 then quantised to i16 with the largest precision that keeps the maximum below 2^15.
static double lanczos3(double x) {
    auto sinc = [](double v) {
        if (v == 0.0) return 1.0;
        v *= 3.14159265358979323846;
        return std::sin(v) / v;
    };
    return (x >= -3.0 && x < 3.0) ? sinc(x) * sinc(x / 3.0) : 0.0;
}

static void build_table(uint32_t in_size, CoefTable& t) {
    const uint32_t out_size = VDF_DCT_SIZE;
    const double scale = (double)in_size / (double)out_size;
    const double fscale = scale > 1.0 ? scale : 1.0;
    const double radius = 3.0 * fscale, recip = 1.0 / fscale;
    const uint32_t window = (uint32_t)std::ceil(radius) * 2 + 1;
    std::vector<double> w((size_t)window * out_size, 0.0);
    t.h_bounds.assign(2 * out_size, 0);
    for (uint32_t o = 0; o < out_size; ++o) {
        const double in_center = ((double)o + 0.5) * scale;
        const uint32_t x_min = (uint32_t)std::fmax(std::floor(in_center - radius), 0.0);
        const uint32_t x_max = (uint32_t)std::fmin(std::ceil(in_center + radius), (double)in_size);
        const double center = in_center - 0.5;
        double* row = w.data() + (size_t)o * window;
        uint32_t start = x_min, end = x_max, cnt = 0;
        double sum = 0.0;
        for (uint32_t x = x_min; x < x_max; ++x) {
            const double v = lanczos3(((double)x - center) * recip);
            if (x == start && v == 0.0) {
                ++start;
                continue;
            }
            row[cnt++] = v;
            sum += v;
        }
        for (uint32_t q = cnt; q > 0 && end > start && row[q - 1] == 0.0; --q) --end;
        if (sum != 0.0)
            for (uint32_t q = 0; q < cnt; ++q) row[q] /= sum;
        t.h_bounds[2 * o] = start;
        t.h_bounds[2 * o + 1] = end - start;
    }
    double max_w = 0.0;
    for (double v : w) max_w = std::fmax(max_w, v);
    uint32_t precision = 0;
    for (uint32_t cur = 0; cur < 22; ++cur) {
        precision = cur;
        const double nv = std::round(max_w * (double)(1 << (cur + 1)));
        if (nv >= (double)(1 << 15)) break;
    }
    t.h_k.resize(w.size());
    const double sc = (double)(1 << precision);
    for (size_t q = 0; q < w.size(); ++q) {
        double v = std::round(w[q] * sc);
        v = std::fmin(std::fmax(v, -32768.0), 32767.0);
        t.h_k[q] = (int16_t)v;
    }
    t.in_size = in_size;
    t.window = window;
    t.precision = precision;
}

static int ensure_luts(vdf_ctx* ctx) {
    if (ctx->h_coef_lut.p) return VDF_OK;
    VDF_ALLOC(ctx, ctx->h_coef_lut.ensure((size_t)(kMaxDim + 1) * sizeof(CoefRef)));
    VDF_ALLOC(ctx, ctx->h_bfrag_lut.ensure((size_t)(kMaxDim + 1) * 16 * sizeof(BFragRef)));
    VDF_CUDA(ctx, cudaMemsetAsync(ctx->h_coef_lut.p, 0, (size_t)(kMaxDim + 1) * sizeof(CoefRef), ctx->stream));
    VDF_CUDA(ctx, cudaMemsetAsync(ctx->h_bfrag_lut.p, 0, (size_t)(kMaxDim + 1) * 16 * sizeof(BFragRef), ctx->stream));
    return VDF_OK;
}

static int get_table(vdf_ctx* ctx, uint32_t in_size, const CoefTable** out) {
    auto it = ctx->coef_cache.find(in_size);
    if (it == ctx->coef_cache.end()) {
        VDF_TRY(ensure_luts(ctx));
        CoefTable t;
        build_table(in_size, t);
        VDF_ALLOC(ctx, cudaMalloc(&t.d_bounds, t.h_bounds.size() * 4));
        VDF_ALLOC(ctx, cudaMalloc(&t.d_k, t.h_k.size() * 2));
        VDF_CUDA(ctx, cudaMemcpyAsync(t.d_bounds, t.h_bounds.data(), t.h_bounds.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaMemcpyAsync(t.d_k, t.h_k.data(), t.h_k.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
        // the same table as A fragments of mma.m16n8k32 for the fused kernel's vertical pass: [32-row k-step][hi, lo][lane] -> (a0..a3);
        // a0 = rows g, k 4q..4q+3; a1 = rows g+8; a2, a3 = the same at k + 16 (g = lane / 4, q = lane % 4); zero outside the window,
        // k-steps padded to whole 128-row blocks
        {
            const uint32_t n_vk = (in_size + 127) / 128 * 4;
            std::vector<uint32_t> va((size_t)n_vk * 2 * 32 * 4);
            for (uint32_t vk = 0; vk < n_vk; ++vk)
                for (uint32_t hl = 0; hl < 2; ++hl)
                    for (uint32_t lane = 0; lane < 32; ++lane)
                        for (uint32_t reg = 0; reg < 4; ++reg) {
                            const uint32_t o = lane / 4 + (reg & 1) * 8, kbase = vk * 32 + (reg >> 1) * 16 + (lane % 4) * 4;
                            const uint32_t start = t.h_bounds[2 * o], size = t.h_bounds[2 * o + 1];
                            uint32_t word = 0;
                            for (uint32_t i = 0; i < 4; ++i) {
                                const uint32_t y = kbase + i;
                                int k = 0;
                                if (y >= start && y < start + size) k = t.h_k[(size_t)o * t.window + (y - start)];
                                word |= (hl == 0 ? (uint32_t)((k >> 8) & 0xFF) : (uint32_t)(k & 0xFF)) << (8 * i);
                            }
                            va[(((size_t)vk * 2 + hl) * 32 + lane) * 4 + reg] = word;
                        }
            VDF_ALLOC(ctx, cudaMalloc(&t.d_kva, va.size() * 4));
            VDF_CUDA(ctx, cudaMemcpyAsync(t.d_kva, va.data(), va.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `va` goes out of scope
        }
        const CoefRef ref{t.d_bounds, t.d_k, t.window, t.precision, reinterpret_cast<const uint4*>(t.d_kva)};
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->h_coef_lut.as<CoefRef>() + in_size, &ref, sizeof ref, cudaMemcpyHostToDevice, ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors may move when the map rebalances
        it = ctx->coef_cache.emplace(in_size, std::move(t)).first;
    }
    *out = &it->second;
    return VDF_OK;
}

void free_coef_cache(vdf_ctx* ctx) {
    for (auto& kv : ctx->coef_cache) {
        if (kv.second.d_bounds) cudaFree(kv.second.d_bounds);
        if (kv.second.d_k) cudaFree(kv.second.d_k);
        if (kv.second.d_kb) cudaFree(kv.second.d_kb);
        if (kv.second.d_kva) cudaFree(kv.second.d_kva);
    }
    ctx->coef_cache.clear();
    for (auto& kv : ctx->bfrag_cache) cudaFree(kv.second);
    ctx->bfrag_cache.clear();
    ctx->h_coef_lut.release();
    ctx->h_bfrag_lut.release();
}

// IMMA B fragments of the horizontal coefficients for a crop that starts `shift` bytes after a 16-byte boundary:
// [k-step][n-tile][lane] -> (b0, b1); n-tiles 0,1 = high bytes (signed) of outputs 0-7 / 8-15, 2,3 = low bytes.
// mma.m16n8k32 B layout: b0 holds k = 4*(lane%4)+0..3, b1 the same +16, column n = lane/4.
static int get_bfrags(vdf_ctx* ctx, const CoefTable& t, uint32_t shift, const uint2** out, const uint8_t** mask_out) {
    const bool dense = ctx->hash_variant == 3;  // experiment knob: never skip coefficient tiles
    const uint64_t key = ((uint64_t)t.in_size << 8) | shift | (dense ? 1ull << 62 : 0);
    auto it = ctx->bfrag_cache.find(key);
    if (it == ctx->bfrag_cache.end()) {
        const uint32_t n_kch = (shift + t.in_size + kKch - 1) / kKch, ksteps = n_kch * 4;
        std::vector<uint32_t> frag((size_t)ksteps * 4 * 32 * 2);
        for (uint32_t ks = 0; ks < ksteps; ++ks)
            for (uint32_t nt = 0; nt < 4; ++nt)
                for (uint32_t lane = 0; lane < 32; ++lane)
                    for (uint32_t reg = 0; reg < 2; ++reg) {
                        uint32_t word = 0;
                        const uint32_t o = (nt & 1) * 8 + lane / 4;
                        const uint32_t start = t.h_bounds[2 * o], size = t.h_bounds[2 * o + 1];
                        for (uint32_t i = 0; i < 4; ++i) {
                            const int64_t x = (int64_t)ks * 32 + (lane % 4) * 4 + i + reg * 16 - shift;
                            int k = 0;
                            if (x >= (int64_t)start && x < (int64_t)start + size) k = t.h_k[(size_t)o * t.window + (x - start)];
                            const uint32_t byte = nt < 2 ? (uint32_t)((k >> 8) & 0xFF) : (uint32_t)(k & 0xFF);
                            word |= byte << (8 * i);
                        }
                        frag[(((size_t)ks * 4 + nt) * 32 + lane) * 2 + reg] = word;
                    }
        // tap masks, appended behind the fragments: per k-chunk bit (2*ks + octet)
        const size_t frag_words = frag.size();
        frag.resize(frag_words + (n_kch + 3) / 4, 0u);
        uint8_t* masks = reinterpret_cast<uint8_t*>(frag.data() + frag_words);
        for (uint32_t ks = 0; ks < ksteps; ++ks)
            for (uint32_t oct = 0; oct < 2; ++oct) {
                bool any = false;
                for (uint32_t lane = 0; lane < 32 && !any; ++lane)
                    for (uint32_t reg = 0; reg < 2 && !any; ++reg)
                        any = frag[(((size_t)ks * 4 + oct) * 32 + lane) * 2 + reg] != 0 ||
                              frag[(((size_t)ks * 4 + 2 + oct) * 32 + lane) * 2 + reg] != 0;
                if (any || dense) masks[ks / 4] |= (uint8_t)(1u << (2 * (ks % 4) + oct));
            }
        // the fused kernel's form, behind the masks (16-byte aligned): per 256-pixel chunk the fragments of its two 128-pixel halves
        // (zeros beyond the last one) and a 16-byte header {mask of half 0, mask of half 1, 0...}
        const std::vector<uint8_t> mask_copy(masks, masks + n_kch);  // `masks` points into `frag`, which is about to grow
        const size_t kb2_word0 = (frag.size() + 3) / 4 * 4, n_kc2 = (n_kch + 1) / 2;
        frag.resize(kb2_word0 + n_kc2 * (kFCoefBytes / 4), 0u);
        for (size_t c = 0; c < n_kc2; ++c) {
            uint32_t* dst = frag.data() + kb2_word0 + c * (kFCoefBytes / 4);
            for (uint32_t h = 0; h < 2; ++h) {
                const size_t kc = 2 * c + h;
                if (kc >= n_kch) continue;
                std::memcpy(dst + h * (kBFragBytes / 4), frag.data() + kc * (kBFragBytes / 4), kBFragBytes);
                reinterpret_cast<uint8_t*>(dst + 2 * (kBFragBytes / 4))[h] = mask_copy[kc];
            }
        }
        void* d = nullptr;
        VDF_ALLOC(ctx, cudaMalloc(&d, frag.size() * 4));
        VDF_CUDA(ctx, cudaMemcpyAsync(d, frag.data(), frag.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        const BFragRef ref{reinterpret_cast<const uint2*>(d), reinterpret_cast<const uint8_t*>(d) + (size_t)n_kch * kBFragBytes,
                           reinterpret_cast<const uint8_t*>(d) + kb2_word0 * 4};
        VDF_TRY(ensure_luts(ctx));
        VDF_CUDA(ctx, cudaMemcpyAsync(ctx->h_bfrag_lut.as<BFragRef>() + (size_t)t.in_size * 16 + shift, &ref, sizeof ref, cudaMemcpyHostToDevice,
                                      ctx->stream));
        VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->bfrag_cache.emplace(key, d).first;
    }
    *out = reinterpret_cast<const uint2*>(it->second);
    *mask_out = reinterpret_cast<const uint8_t*>(it->second) + (size_t)((shift + t.in_size + kKch - 1) / kKch) * kBFragBytes;
    return VDF_OK;
}

static int load_dct_consts(vdf_ctx* ctx) {
    if (ctx->dct_consts_loaded) return VDF_OK;
    // rustdct::twiddles::single_twiddle(i, fft_len).conj()
    auto tw = [](unsigned i, unsigned fft_len, double* re, double* im) {
        const double constant = -2.0 * 3.14159265358979323846 / (double)fft_len;
        const double angle = constant * (double)i;
        *re = std::cos(angle);
        *im = -std::sin(angle);
    };
    double h[16] = {0};
    for (unsigned i = 0; i < 4; ++i) tw(2 * i + 1, 64, &h[2 * i], &h[2 * i + 1]);
    for (unsigned i = 0; i < 2; ++i) tw(2 * i + 1, 32, &h[8 + 2 * i], &h[8 + 2 * i + 1]);
    tw(1, 16, &h[12], &h[13]);
    h[14] = 0.70710678118654752440;  // std::f64::consts::FRAC_1_SQRT_2
    VDF_CUDA(ctx, cudaMemcpyToSymbolAsync(c_tw, h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    VDF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->dct_consts_loaded = true;
    return VDF_OK;
}

int hash_from_small_device(vdf_ctx* ctx, const uint8_t* d_small, uint32_t n, uint64_t* d_out_hash) {
    if (n == 0) return VDF_OK;
    VDF_TRY(load_dct_consts(ctx));
    dct_pack_kernel<<<n, 256, 0, ctx->stream>>>(d_small, nullptr, n, reinterpret_cast<uint32_t*>(d_out_hash));
    VDF_LAUNCHED(ctx);
    return VDF_OK;
}

int hash_stacks_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                       uint64_t* d_out_hash, uint8_t* d_out_small, int32_t* out_status, uint32_t* out_crop) {
    if (n == 0) return VDF_OK;
    if (cropdetect != VDF_CROPDETECT_NONE && cropdetect != VDF_CROPDETECT_LETTERBOX && cropdetect != VDF_CROPDETECT_MOTION) {
        ctx->err = "cropdetect must be VDF_CROPDETECT_NONE, VDF_CROPDETECT_LETTERBOX or VDF_CROPDETECT_MOTION";
        return VDF_ERR_INVALID;
    }
    VDF_TRY(load_dct_consts(ctx));
    VDF_TRY(ensure_luts(ctx));
    cudaStream_t st = ctx->stream;
    if (!ctx->lb_stream) {
        VDF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->lb_stream, cudaStreamNonBlocking));
        VDF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
    }
    const bool four_warps = ctx->hash_variant != 2;  // default: 4 warps, two CTAs per SM
    const size_t ring = four_warps ? ResizeMma<4>::kRingBytes : ResizeMma<8>::kRingBytes;
    const bool allow_fast = ctx->hash_variant != 1;
    const bool use_fused = allow_fast && ctx->hash_variant == 0 && ctx->hash_fused != 0;  // the fused kernel has no size limits of its own
    // status per stack, decided on the host exactly where the reference decides it; tables of the UNCROPPED sizes are made
    // sure of here (most stacks have no bars), every other size is met through a miss
    VDF_ALLOC(ctx, ctx->pin_b.ensure((size_t)n * (sizeof(StackDev) + 4 + 4)));
    StackDev* sd = ctx->pin_b.as<StackDev>();
    int32_t* status = reinterpret_cast<int32_t*>(sd + n);
    uint32_t* h_miss = reinterpret_cast<uint32_t*>(status + n);
    uint32_t max_h = 1;
    bool any_unaligned = false, all_fit = true;
    uint64_t last_wh = ~0ull;
    for (uint32_t s = 0; s < n; ++s) {
        const vdf_stack_desc& d = desc[s];
        int32_t stt = VDF_STACK_OK;
        if (d.flags & VDF_STACK_FLAG_MIXED_SIZES) stt = VDF_STACK_VIDPROC;         // video_hash_builder.rs:169-186
        else if (d.n_frames < VDF_DCT_SIZE) stt = VDF_STACK_NOT_ENOUGH_FRAMES;     // dct_3d.rs:47-52
        else if (d.width == 0 || d.height == 0 || d.pitch < d.width || d.width > kMaxDim || d.height > kMaxDim) {
            ctx->err = "stack " + std::to_string(s) + ": bad geometry (frames up to " + std::to_string(kMaxDim) + " pixels a side)";
            return VDF_ERR_INVALID;
        }
        status[s] = stt;
        // tensor-core path needs 16-byte aligned rows (cp.async 16 B): base, frame stride and pitch
        const bool aligned = ((reinterpret_cast<uintptr_t>(d_frames) + d.offset) % 16 == 0) && d.frame_stride % 16 == 0 && d.pitch % 16 == 0;
        sd[s] = StackDev{d.offset, d.frame_stride, d.width, d.height, d.pitch, stt, aligned ? 1u : 0u, 0u};
        if (stt != VDF_STACK_OK) continue;
        max_h = std::max(max_h, d.height);
        any_unaligned |= !aligned;
        const uint64_t wh = ((uint64_t)d.width << 32) | d.height;
        if (wh != last_wh) {
            const CoefTable *th, *tv;
            VDF_TRY(get_table(ctx, d.width, &th));
            VDF_TRY(get_table(ctx, d.height, &tv));
            const bool fits = (15 + d.width + kKch - 1) / kKch <= 256 && (size_t)32 * tv->window <= ring;  // a crop only shrinks both
            all_fit &= fits;
            if (aligned && (fits || use_fused) && allow_fast) {
                const uint2* kb;
                const uint8_t* km;
                VDF_TRY(get_bfrags(ctx, *th, 0, &kb, &km));
            }
            last_wh = wh;
        }
    }
    const size_t tmp_bytes = (size_t)max_h * 16;
    if (tmp_bytes + ring > 220 * 1024 && !(use_fused && !any_unaligned)) {
        ctx->err = "frame height beyond the resize kernels' shared-memory budget";
        return VDF_ERR_INVALID;
    }
    VDF_ALLOC(ctx, ctx->pin_a.ensure(((size_t)n * 5 + 16) * 4));  // crops (+ in fused mode: miss flags, control words) as they come back
    uint32_t* crop = ctx->pin_a.as<uint32_t>();
    VDF_ALLOC(ctx, ctx->h_jobs.ensure((size_t)n * sizeof(StackJob)));
    VDF_ALLOC(ctx, ctx->h_desc.ensure((size_t)n * sizeof(StackDev)));
    VDF_ALLOC(ctx, ctx->h_sides.ensure((size_t)n * 8 * 4));
    VDF_ALLOC(ctx, ctx->h_crop.ensure((size_t)n * 4 * 4));
    VDF_ALLOC(ctx, ctx->h_done.ensure((size_t)n * 4 * 2 + 64));
    StackJob* d_jobs = ctx->h_jobs.as<StackJob>();
    StackDev* d_sd = ctx->h_desc.as<StackDev>();
    uint32_t* d_done = ctx->h_done.as<uint32_t>();
    uint32_t* d_miss = d_done + n;
    uint32_t* d_nmiss = d_miss + n;
    uint32_t* d_crop = ctx->h_crop.as<uint32_t>();
    uint32_t* d_ctl = nullptr;
    if (use_fused) {
        // one buffer, so that a call is: descriptors up, ONE memset, ONE launch, ONE read-back.  [crop 4n | miss n] survive a second pass;
        // [ctl 16 | flags n | sides_done n | dq n | done n | wq 8n] are zeroed before every launch; ctl[15] counts the misses
        VDF_ALLOC(ctx, ctx->h_fctl.ensure((16 + 17 * (size_t)n) * 4));
        d_crop = ctx->h_fctl.as<uint32_t>();
        d_miss = d_crop + 4 * (size_t)n;
        d_ctl = d_miss + n;
        d_nmiss = d_ctl + 15;
        d_done = d_ctl + 16 + 3 * (size_t)n;
    }
    uint8_t* d_small = d_out_small;
    if (!d_small) {
        VDF_ALLOC(ctx, ctx->h_small.ensure((size_t)n * 4096));
        d_small = ctx->h_small.as<uint8_t>();
    }
    uint32_t* d_hash32 = ctx->hash_fuse_dct ? reinterpret_cast<uint32_t*>(d_out_hash) : nullptr;
    bool any_bad = false;
    for (uint32_t s = 0; s < n && !any_bad; ++s) any_bad = status[s] != VDF_STACK_OK;
    const bool fuse = ctx->hash_fuse_dct != 0;  // DCT + pack in the resize kernel's last CTA per stack, or as a kernel of its own
    VDF_CUDA(ctx, cudaMemcpyAsync(d_sd, sd, (size_t)n * sizeof(StackDev), cudaMemcpyHostToDevice, st));
    if (!use_fused) VDF_CUDA(ctx, cudaMemsetAsync(d_done, 0, (size_t)n * 8 + 64, st));
    if (d_out_hash && any_bad && fuse) VDF_CUDA(ctx, cudaMemsetAsync(d_out_hash, 0, (size_t)n * 128, st));  // stacks that are not hashed read as zero
    // The opt-in is per device AND per function, whoever sets it last wins: every context asks for the device's maximum, once,
    // so that contexts sharing a GPU can never lower each other's limit (kept per context because it is per device).
    const size_t gen_smem = std::max<size_t>(tmp_bytes, kCubeBytes);
    if (!ctx->hash_smem_set[0]) {
        int optin = 0;
        VDF_CUDA(ctx, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        const void* fns[3] = {(const void*)resize_mma_kernel<4>, (const void*)resize_mma_kernel<8>, (const void*)resize_general_kernel};
        size_t least = (size_t)optin;
        for (const void* fn : fns) {  // dynamic limit = opt-in maximum minus the kernel's static shared memory
            cudaFuncAttributes fa;
            VDF_CUDA(ctx, cudaFuncGetAttributes(&fa, fn));
            const size_t dyn = (size_t)optin - fa.sharedSizeBytes;
            VDF_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            least = std::min(least, dyn);
        }
        ctx->hash_smem_set[0] = least;
    }
    if ((tmp_bytes + ring > ctx->hash_smem_set[0] || gen_smem > ctx->hash_smem_set[0]) && !(use_fused && !any_unaligned)) {
        ctx->err = "frame height beyond the resize kernels' shared-memory budget";
        return VDF_ERR_INVALID;
    }
    const bool run_general = any_unaligned || !all_fit || !allow_fast;

    // Chunks of stacks: the letterbox scan of chunk k+1 runs on a second stream beside the resize of chunk k (the scan is a
    // latency-bound strip walk over 0.6 % of the bytes: alone it held the GPU for 12 % of the step in round 1)
    const uint32_t kMaxChunks = 4;
    const uint32_t n_chunks = std::max(1u, std::min(std::min(kMaxChunks, ctx->hash_chunks), n / 32));
    auto chunk_begin = [&](uint32_t k) { return (uint32_t)((uint64_t)n * k / n_chunks); };
    const bool letterbox = cropdetect == VDF_CROPDETECT_LETTERBOX;
    cudaStream_t lb = ctx->hash_overlap ? ctx->lb_stream : st;
    if (letterbox && use_fused) {
        // the fused kernel scans
    } else if (letterbox) {
        if (any_bad) VDF_CUDA(ctx, cudaMemsetAsync(ctx->h_sides.p, 0, (size_t)n * 8 * 4, st));  // the scan writes every side of every good stack
        if (lb != st) {
            VDF_CUDA(ctx, cudaEventRecord(ctx->ev_in, st));
            VDF_CUDA(ctx, cudaStreamWaitEvent(lb, ctx->ev_in, 0));
        }
        kt_begin(ctx, 2, lb);
        if (lb != st) {  // all scans are enqueued up front on their own stream; the resize of chunk k waits for scan k only
            for (uint32_t k = 0; k < n_chunks; ++k) {
                const uint32_t s0 = chunk_begin(k), cnt = chunk_begin(k + 1) - s0;
                VDF_TRY(letterbox_scan(ctx, lb, d_frames, d_sd + s0, cnt, 2u, 8u, nullptr, ctx->h_sides.as<uint32_t>() + (size_t)s0 * 8));
                crop_combine_kernel<<<(cnt + 127) / 128, 128, 0, lb>>>(d_sd + s0, ctx->h_sides.as<uint32_t>() + (size_t)s0 * 8, cnt, 2u,
                                                                      ctx->h_crop.as<uint32_t>() + (size_t)s0 * 4);
                VDF_LAUNCHED(ctx);
                VDF_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[k], lb));
            }
            kt_end(ctx, 2, lb);
        }
    } else if (cropdetect == VDF_CROPDETECT_MOTION) {
        VDF_TRY(motion_crop_device(ctx, d_frames, d_sd, sd, n, d_crop));  // Cropdetect::Motion (motion.cu)
    } else {
        VDF_CUDA(ctx, cudaMemsetAsync(d_crop, 0, (size_t)n * 16, st));  // Cropdetect::None: zero crop (:195-199)
    }
    // ---- default: ONE persistent launch per pass (hash_fused_kernel) does the scan, the crops, the jobs, the resize and the DCT
    if (use_fused) {
        if (!ctx->hash_smem_set[1]) {
            VDF_CUDA(ctx, cudaFuncSetAttribute((const void*)hash_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes));
            ctx->hash_smem_set[1] = kFusedSmemBytes;
        }
    }
    auto run_fused = [&](bool only_missed, bool scan_in_kernel) -> int {
        VDF_CUDA(ctx, cudaMemsetAsync(d_ctl, 0, (16 + (scan_in_kernel ? 12 : 4) * (size_t)n) * 4, st));
        FusedArgs fa;
        fa.frames = d_frames, fa.stacks = d_sd, fa.jobs = d_jobs, fa.n = n, fa.n_lb_items = scan_in_kernel ? n * 8u : 0u;
        fa.sides = ctx->h_sides.as<uint32_t>(), fa.crop = d_crop;
        fa.coef_lut = ctx->h_coef_lut.as<CoefRef>(), fa.bfrag_lut = ctx->h_bfrag_lut.as<BFragRef>();
        fa.miss = d_miss, fa.n_miss = d_nmiss;
        fa.ctl = d_ctl, fa.flags = d_ctl + 16, fa.sides_done = d_ctl + 16 + n, fa.dq = d_ctl + 16 + 2 * (size_t)n, fa.wq = d_ctl + 16 + 4 * (size_t)n;
        fa.done = d_done, fa.small = d_small, fa.out_hash = d_hash32;
        fa.exp = 0u;
#ifdef VDF_TIMING_EXPERIMENTS
        if (getenv("VDF_FUSED_EXP")) fa.exp = (uint32_t)atoi(getenv("VDF_FUSED_EXP"));
#endif
        if (!only_missed) kt_begin(ctx, 1);
        if (!scan_in_kernel) {  // crops known up front (Cropdetect::None / Motion, or the pass for the sizes met for the first time)
            job_build_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_sd, d_crop, n, fa.coef_lut, fa.bfrag_lut, 0xFFFFFFFFu, 0xFFFFFFFFu,
                                                              1u, only_missed ? 1u : 0u, d_jobs, d_miss, d_nmiss, fa.flags, fa.ctl + 5);
            VDF_LAUNCHED(ctx);
        }
        hash_fused_kernel<<<(unsigned)ctx->sm_count, kFThreads, kFusedSmemBytes, st>>>(fa);
        VDF_LAUNCHED(ctx);
        if (!only_missed) kt_end(ctx, 1);
        if (any_unaligned) {  // stacks whose rows are not 16-byte aligned: the general kernel (it skips every job that is `fast`)
            resize_general_kernel<<<n * 16, 256, gen_smem, st>>>(d_frames, d_jobs, d_small, d_done, d_hash32);
            VDF_LAUNCHED(ctx);
        }
        return VDF_OK;
    };
    auto run_pass = [&](bool only_missed) -> int {
        if (use_fused) return run_fused(only_missed, letterbox && !only_missed);
        for (uint32_t k = 0; k < n_chunks; ++k) {
            const uint32_t s0 = chunk_begin(k), cnt = chunk_begin(k + 1) - s0;
            if (letterbox && !only_missed) {
                if (lb != st) {
                    VDF_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_chunk[k], 0));
                } else {
                    VDF_TRY(letterbox_scan(ctx, st, d_frames, d_sd + s0, cnt, 2u, 8u, nullptr, ctx->h_sides.as<uint32_t>() + (size_t)s0 * 8));
                    crop_combine_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(d_sd + s0, ctx->h_sides.as<uint32_t>() + (size_t)s0 * 8, cnt, 2u,
                                                                          ctx->h_crop.as<uint32_t>() + (size_t)s0 * 4);
                    VDF_LAUNCHED(ctx);
                    if (k == n_chunks - 1) kt_end(ctx, 2);
                }
            }
            if (k == 0 && !only_missed) kt_begin(ctx, 1);
            job_build_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(d_sd + s0, ctx->h_crop.as<uint32_t>() + (size_t)s0 * 4, cnt,
                                                                ctx->h_coef_lut.as<CoefRef>(), ctx->h_bfrag_lut.as<BFragRef>(), (uint32_t)ring, 256u,
                                                                allow_fast ? 1u : 0u, only_missed ? 1u : 0u, d_jobs + s0, d_miss + s0, d_nmiss,
                                                                nullptr, nullptr);
            VDF_LAUNCHED(ctx);
            if (allow_fast) {
                if (four_warps)
                    resize_mma_kernel<4><<<cnt * 16, 128, tmp_bytes + ring, st>>>(d_frames, d_jobs + s0, d_small + (size_t)s0 * 4096, d_done + s0,
                                                                                 d_hash32 ? d_hash32 + (size_t)s0 * 32 : nullptr);
                else
                    resize_mma_kernel<8><<<cnt * 16, 256, tmp_bytes + ring, st>>>(d_frames, d_jobs + s0, d_small + (size_t)s0 * 4096, d_done + s0,
                                                                                 d_hash32 ? d_hash32 + (size_t)s0 * 32 : nullptr);
                VDF_LAUNCHED(ctx);
            }
            if (run_general || only_missed) {
                resize_general_kernel<<<cnt * 16, 256, gen_smem, st>>>(d_frames, d_jobs + s0, d_small + (size_t)s0 * 4096, d_done + s0,
                                                                       d_hash32 ? d_hash32 + (size_t)s0 * 32 : nullptr);
                VDF_LAUNCHED(ctx);
            }
        }
        if (!only_missed) kt_end(ctx, 1);
        return VDF_OK;
    };
    VDF_TRY(run_pass(false));
    auto separate_dct = [&]() -> int {
        if (fuse || !d_out_hash) return VDF_OK;
        VDF_ALLOC(ctx, ctx->h_hash.ensure(16));
        // status per stack on the device for the stand-alone kernel: the jobs carry it (kJobSkip = missed in this pass)
        kt_begin(ctx, 3);
        dct_pack_jobs_kernel<<<n, 256, 0, st>>>(d_small, d_jobs, n, reinterpret_cast<uint32_t*>(d_out_hash));
        kt_end(ctx, 3);
        VDF_LAUNCHED(ctx);
        return VDF_OK;
    };
    VDF_TRY(separate_dct());
    VDF_ALLOC(ctx, ctx->h_misc.ensure(256));
    uint32_t* h_n = reinterpret_cast<uint32_t*>(ctx->h_misc.as<unsigned long long>() + 24);
    if (use_fused) {  // crops, miss flags and the control words in one copy
        VDF_CUDA(ctx, cudaMemcpyAsync(crop, d_crop, ((size_t)n * 5 + 16) * 4, cudaMemcpyDeviceToHost, st));
        h_miss = crop + 4 * (size_t)n;
        h_n = crop + 5 * (size_t)n + 15;
    } else {
        VDF_CUDA(ctx, cudaMemcpyAsync(crop, d_crop, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        VDF_CUDA(ctx, cudaMemcpyAsync(h_n, d_nmiss, 4, cudaMemcpyDeviceToHost, st));
    }
    const uint32_t* h_ctl = crop + 5 * (size_t)n;  // fused mode only
    VDF_CUDA(ctx, cudaStreamSynchronize(st));
    if (use_fused && h_ctl[6]) {
        ctx->err = "hash_fused_kernel: a wait timed out (code " + std::to_string(h_ctl[6]) + ")";
        return VDF_ERR_CUDA;
    }
    if (use_fused && getenv("VDF_FUSED_STATS")) {  // the kernel's own wait statistics (~us, summed over blocks)
        fprintf(stderr, "fused stats: frames %u lb_items %u/%u dct %u/%u nofuse %u | wait_job_us %u wait_first_slot_us %u wait_slot_us %u scan_end_us %u pixel_life_us_sum %u\n", h_ctl[0],
                h_ctl[1], h_ctl[2], h_ctl[3], h_ctl[4], h_ctl[5], h_ctl[8], h_ctl[14], h_ctl[9], h_ctl[10], h_ctl[11]);
    }
    if (*h_n) {  // sizes met for the first time: build their tables (host, f64 + libm sin like the reference), run those stacks
        if (!use_fused) {
            VDF_CUDA(ctx, cudaMemcpyAsync(h_miss, d_miss, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            VDF_CUDA(ctx, cudaStreamSynchronize(st));
        }
        for (uint32_t s = 0; s < n; ++s) {
            if (!h_miss[s]) continue;
            const uint32_t* c = &crop[(size_t)s * 4];
            const uint32_t cw = desc[s].width - c[0] - c[1], chh = desc[s].height - c[2] - c[3];
            const CoefTable *th, *tv;
            VDF_TRY(get_table(ctx, cw, &th));
            VDF_TRY(get_table(ctx, chh, &tv));
            const uint32_t shift = c[0] & 15u;
            if (sd[s].aligned && allow_fast && (use_fused || ((shift + cw + kKch - 1) / kKch <= 256 && (size_t)32 * tv->window <= ring))) {
                const uint2* kb;
                const uint8_t* km;
                VDF_TRY(get_bfrags(ctx, *th, shift, &kb, &km));
            }
        }
        VDF_CUDA(ctx, cudaMemsetAsync(d_nmiss, 0, 4, st));
        VDF_TRY(run_pass(true));
        VDF_TRY(separate_dct());
        if (use_fused) VDF_CUDA(ctx, cudaMemcpyAsync(crop + 5 * (size_t)n, d_ctl, 64, cudaMemcpyDeviceToHost, st));
        else VDF_CUDA(ctx, cudaMemcpyAsync(h_n, d_nmiss, 4, cudaMemcpyDeviceToHost, st));
        VDF_CUDA(ctx, cudaStreamSynchronize(st));
        if (use_fused && h_ctl[6]) {
            ctx->err = "hash_fused_kernel: a wait timed out (code " + std::to_string(h_ctl[6]) + ")";
            return VDF_ERR_CUDA;
        }
        if (*h_n) {
            ctx->err = "resize tables still missing after the second pass";
            return VDF_ERR_CUDA;
        }
    }
    if (out_status) std::memcpy(out_status, status, (size_t)n * 4);
    if (out_crop) std::memcpy(out_crop, crop, (size_t)n * 16);
    return VDF_OK;
}

}  // namespace vdf
