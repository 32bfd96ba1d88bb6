/*
 * vdf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the two hot paths of Farmadupe/vid_dup_finder_lib:
 *   (1) frame stack -> VideoHash    (letterbox crop, Lanczos3 u8 resize, 16^3 f64 DCT-II, bit pack)
 *   (2) Hamming search              (search / search_with_references -> MatchGroups)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / CPU baseline.  The product (libvdf_b200.so) never
 * links, loads or calls anything in oracle/.
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   search path   : PINNED  - re-expresses every search/Hamming test the reference holds
 *                             (tests/test_oracle_search.py <- vid_dup_finder_lib/tests/test_find_all.rs,
 *                             search_algorithm.rs:203-208, video_hash.rs:325-371).
 *   letterbox/crop: PINNED  - 17 exact Crop KATs (video_frames_gray.rs:225-458), crop.rs:204-365.
 *   resize + DCT  : PARITY UNPINNED - the arithmetic lives in the un-vendored crates
 *                             fast_image_resize 5.1 and rustdct 0.7 (no Cargo.lock, no source under
 *                             /root/reference, no golden hash values in any reference test).  The
 *                             restatement follows their published algorithms; it is cross-checked
 *                             against Pillow LANCZOS and scipy.fft.dctn and against the reference's
 *                             example clips (groups form as examples/example.rs:77-82 expects).
 *
 * All citations are file:line relative to /root/reference.
 */
#ifndef VDF_ORACLE_H
#define VDF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDFO_DCT_SIZE 16    /* definitions.rs:34 */
#define VDFO_HASH_SIZE 10   /* definitions.rs:36 */
#define VDFO_HASH_BITS 1000 /* definitions.rs:42 */
#define VDFO_HASH_WORDS 16  /* definitions.rs:43 (usize = 64 bit) */

/* status codes mirror vid_dup_finder_lib/src/video_hashing/mod.rs:17-28 */
#define VDFO_OK 0
#define VDFO_NOT_ENOUGH_FRAMES 1
#define VDFO_VIDPROC 2

/* ---------------------------------------------------------------- search path */

/* video_hash.rs:311-317 : sum over all 16 words of popcount(x^y) (1024 bits, pad bits included) */
uint32_t vdfo_hamming(const uint64_t* x, const uint64_t* y);

/* search_algorithm.rs:82 : (tolerance * TOLERANCE_SCALING_FACTOR) as u32  (Rust saturating cast) */
uint32_t vdfo_tolerance_int(double tolerance);

/* search_algorithm.rs:99 : (f64::from(lhs_duration) * 1.1) as u32 */
uint32_t vdfo_self_window_thresh(uint32_t duration);

/* search_algorithm.rs:174,179 : ((d*0.95) as u32, (d*1.05) as u32) */
void vdfo_ref_window_durations(uint32_t duration, uint32_t* lo, uint32_t* hi);

/* Rust std::path::Path::cmp on Unix (component-wise); returns <0, 0, >0 */
int vdfo_path_cmp(const char* a, const char* b);

/* search_algorithm.rs:55-61 : stable sort by (duration, src_path); writes the permutation */
void vdfo_sort_order(const uint32_t* duration, const char* const* paths, uint64_t n, uint64_t* order_out);

/* search_algorithm.rs:81-171 literal: lhs/rhs cursor walk with the matched flags.
 * Inputs are already sorted.  Output CSR: group g = members[group_ptr[g] .. group_ptr[g+1]),
 * indices into the sorted array, matches ascending then the target LAST, groups already reversed. */
int vdfo_search_self(const uint64_t* hashes_sorted, const uint32_t* dur_sorted, uint64_t n, uint32_t tol_int,
                     uint64_t** group_ptr_out, uint64_t** members_out, uint64_t* n_groups_out);

/* brute-force edge list of SURVEY A.2: all i<j with dur[j] <= thresh(dur[i]) and hamming <= tol,
 * sorted by (i,j); edges_out holds 2*n_edges values (i0,j0,i1,j1,...) */
int vdfo_self_edges(const uint64_t* hashes_sorted, const uint32_t* dur_sorted, uint64_t n, uint32_t tol_int,
                    uint64_t** edges_out, uint64_t* n_edges_out);

/* closed form of the greedy rule over an (i,j)-sorted edge list; same CSR as above */
int vdfo_group_from_edges(uint64_t n, const uint64_t* edges, uint64_t n_edges, uint64_t** group_ptr_out,
                          uint64_t** members_out, uint64_t* n_groups_out);

/* video_dup_finder.rs:19-46 + search_algorithm.rs:63-77,173-185 ; CSR row per ref (caller order),
 * col_idx ascending indices into the sorted candidate array */
int vdfo_search_refs(const uint64_t* cand_sorted, const uint32_t* cand_dur_sorted, uint64_t n_cand,
                     const uint64_t* refs, const uint32_t* ref_dur, uint64_t n_ref, uint32_t tol_int,
                     uint64_t** row_ptr_out, uint64_t** col_idx_out);

/* pair counts the reference would evaluate hamming_distance on (for throughput accounting) */
uint64_t vdfo_self_window_pairs(const uint32_t* dur_sorted, uint64_t n);

void vdfo_free(void* p);

/* ---------------------------------------------------------------- hashing path */

#define VDFO_LB_BLACKWHITE 0
#define VDFO_LB_ANYCOLOUR 1

/* video_frames_gray.rs:38-128 ; out = {left,right,top,bottom} */
void vdfo_letterbox_frame(const uint8_t* pix, uint32_t w, uint32_t h, size_t pitch, int colour_mode, uint8_t tol,
                          uint32_t out_lrtb[4]);

/* video_frames_gray.rs:201-210 + crop.rs:53-68 : frames step_by(8).take(8), AnyColour(16), per-side min */
int vdfo_cropdetect_letterbox(const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h, size_t pitch,
                              size_t frame_stride, uint32_t out_lrtb[4]);

/* fast_image_resize 5.1 coefficient generation (convolution/mod.rs precompute_coefficients +
 * optimisations.rs Normalizer16), Lanczos3, adaptive kernel.  bounds_out: out_size pairs (start,size);
 * coefs_out: out_size*window i16; returns window size via *window_out, precision via *precision_out. */
int vdfo_resize_coeffs(uint32_t in_size, uint32_t out_size, uint32_t** bounds_out, int16_t** coefs_out,
                       uint32_t* window_out, uint32_t* precision_out);

/* resize_gray.rs:11-54 : crop window (left,top,cw,ch) of src -> out_w x out_h u8, horizontal pass into a
 * u8 temp then vertical pass */
int vdfo_resize_lanczos3(const uint8_t* src, uint32_t w, uint32_t h, size_t pitch, uint32_t left, uint32_t top,
                         uint32_t cw, uint32_t ch, uint32_t out_w, uint32_t out_h, uint8_t* dst);

/* rustdct 0.7 plan_dct2(16): unnormalised DCT-II, split-radix butterfly, f64, in place */
void vdfo_dct2_16(double* buf);

/* raw_dct_ops.rs:107-142 : cube[t][x][y], DCT along y, then x, then t */
void vdfo_dct3d(double* cube /* 4096 */);

/* dct_3d.rs:15-66 + video_hash.rs:63-70 : small = 16 frames of 16x16 u8 (row-major, [t][row][col]);
 * optional coef_out receives the 4096 f64 coefficients in [t][x][y] order */
void vdfo_hash_from_small(const uint8_t* small, uint64_t hash_out[16], double* coef_out);

/* video_hash_builder.rs:169-212 + video_hash.rs:45-73 : one stack -> hash.  cropdetect 0 None, 1 Letterbox.
 * frame_dims (optional, n_frames pairs w,h) lets a test inject a size mismatch (-> VDFO_VIDPROC). */
int vdfo_hash_stack(const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h, size_t pitch,
                    size_t frame_stride, int cropdetect, const uint32_t* frame_dims, uint64_t hash_out[16],
                    uint32_t crop_out_lrtb[4], uint8_t* small_out /* optional 4096 */);

/* batch helper used by the CPU baseline: OpenMP-free, caller threads it */
int vdfo_hash_stacks(const uint8_t* frames, uint64_t n_stacks, uint32_t w, uint32_t h, int cropdetect,
                     uint64_t* hash_out, int32_t* status_out);

#ifdef __cplusplus
}
#endif
#endif
