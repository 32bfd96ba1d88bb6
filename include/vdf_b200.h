/*
 * vdf_b200.h -- C ABI of the B200-native (sm_100a) replacement for the two data-parallel hot paths of
 * Farmadupe/vid_dup_finder_lib.  Plain pointers and sizes only; this is what the crate's `extern "C"`
 * block (or any other FFI: ctypes, cgo, JNI) binds.  Built by vid_dup_finder_lib_b200/csrc/Makefile into
 * libvdf_b200.so (Python/ctypes, tests, bench) and libvdf_b200.a (the static library build.rs links).
 *
 * The reference has no FFI seam of its own; its seam is the public Rust API
 * (vid_dup_finder_lib/src/lib.rs:132-140).  Each entry point below names the reference code it replaces
 * (file:line relative to the reference repository root).  What stays in the host language above this
 * boundary: decoding, the stable (duration, path) sort (search_algorithm.rs:55-61), the tolerance cast
 * (search_algorithm.rs:82), index -> PathBuf, MatchGroup construction (match_group.rs:21-47).
 *
 * There is NO CPU fallback: every compute entry point runs CUDA kernels and returns VDF_ERR_CUDA /
 * VDF_ERR_NO_DEVICE if it cannot.
 *
 * Conventions
 *   - hashes: 16 x u64 per VideoHash, exactly `VideoHash.hash: [usize; 16]` (video_hash.rs:26-32);
 *     bit b of the 1000-bit hash lives in word b/64, bit b%64; bits 1000..1023 are zero for real hashes
 *     but ARE compared (video_hash.rs:311-317).
 *   - "sorted" inputs are in the order of Search::sort: ascending (duration, src_path).
 *   - indices fit in 32 bits (n < 2^32).
 *   - inputs are caller-owned and read-only for the duration of the call.  `*_device` variants take
 *     device pointers valid on the context's GPU; all others take host pointers (pinned or pageable).
 *   - outputs in vdf_edges / vdf_groups / vdf_csr are library-allocated host buffers, released with
 *     vdf_free_edges / vdf_free_groups / vdf_free_csr.
 *   - return value: VDF_OK (0) or a negative error; vdf_last_error(ctx) describes the last failure.
 *     Nothing unwinds or aborts across the boundary.
 *   - a context is bound to one GPU (vdf_ctx_create; one process per GPU) or to several GPUs of one node
 *     (vdf_ctx_create_multi; one process drives them all), owns its streams and scratch buffers, and is
 *     NOT thread-safe: one context per host thread, or an external lock
 *     (the reference calls VideoHashBuilder::hash from rayon workers,
 *     vid_dup_finder_app/src/video_hash_filesystem_cache/video_hash_filesystem_cache.rs:246).
 */
#ifndef VDF_B200_H
#define VDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDF_HASH_WORDS 16 /* definitions.rs:43 */
#define VDF_DCT_SIZE 16   /* definitions.rs:34 */
#define VDF_HASH_BITS 1000 /* definitions.rs:42 */

#define VDF_OK 0
#define VDF_ERR_CUDA (-1)          /* a CUDA call or kernel failed */
#define VDF_ERR_ALLOC (-2)         /* host or device allocation failed */
#define VDF_ERR_INVALID (-3)       /* bad argument */
#define VDF_ERR_EDGE_OVERFLOW (-4) /* more matches than the edge buffer may grow to (see "max_edges") */
#define VDF_ERR_NO_DEVICE (-5)     /* no usable sm_100 device */
#define VDF_ERR_IO (-6)            /* a cache file could not be opened, written or renamed */
#define VDF_ERR_FORMAT (-7)        /* a cache file is truncated or not in the expected encoding */

/* per-stack status, mirrors vid_dup_finder_lib::Error (video_hashing/mod.rs:17-28) */
#define VDF_STACK_OK 0
#define VDF_STACK_NOT_ENOUGH_FRAMES 1 /* Error::NotEnoughFrames: fewer than 16 frames (dct_3d.rs:47-52) */
#define VDF_STACK_VIDPROC 2           /* Error::VidProc: frames not all the same size (video_hash_builder.rs:169-186) */

#define VDF_CROPDETECT_NONE 0      /* Cropdetect::None      (definitions.rs:46-54) */
#define VDF_CROPDETECT_LETTERBOX 1 /* Cropdetect::Letterbox (the library default, video_hash_builder.rs:55-63) */
#define VDF_CROPDETECT_MOTION 2    /* Cropdetect::Motion: MotiondetectCrop::from_frames (vid_dup_finder_common/src/motioncrop/
                                      autocrop_frames.rs:36-218), csrc/motion.cu */

typedef struct vdf_ctx vdf_ctx;
typedef struct vdf_table vdf_table; /* `Search` (search_algorithm.rs:5-23): a sorted table prepared for searching, resident in HBM */

/* (i,j) match edges of search_self, i<j, sorted by (i,j): ij[2k], ij[2k+1] */
typedef struct {
    uint64_t n;
    uint64_t* ij;
} vdf_edges;

/* MatchGroups as CSR over indices into the sorted array: group g = member_idx[group_ptr[g] .. group_ptr[g+1]),
 * matches in ascending order followed by the target LAST, groups in the reference's (reversed) order
 * (search_algorithm.rs:136,158-161,167) */
typedef struct {
    uint64_t n_groups;
    uint64_t* group_ptr;  /* n_groups + 1 */
    uint64_t* member_idx; /* group_ptr[n_groups] */
} vdf_groups;

/* per-reference match lists: row r (caller order) = col_idx[row_ptr[r] .. row_ptr[r+1]), ascending indices
 * into the sorted candidate array (search_algorithm.rs:63-77) */
typedef struct {
    uint64_t n_rows;
    uint64_t* row_ptr; /* n_rows + 1 */
    uint64_t* col_idx;
} vdf_csr;

/* one frame stack = the <=16 gray u8 frames VideoHashBuilder collects (video_hash_builder.rs:159-167) */
typedef struct {
    uint64_t offset;       /* byte offset of frame 0 from the `frames` base pointer */
    uint64_t frame_stride; /* bytes between consecutive frames of this stack */
    uint32_t width;        /* pixels; every frame of the stack has this size */
    uint32_t height;
    uint32_t pitch;        /* bytes between rows (>= width) */
    uint32_t n_frames;     /* frames present; the first 16 are hashed, fewer -> VDF_STACK_NOT_ENOUGH_FRAMES */
    uint32_t flags;        /* VDF_STACK_FLAG_* */
    uint32_t reserved;
} vdf_stack_desc;
#define VDF_STACK_FLAG_MIXED_SIZES 1u /* host saw differing frame sizes -> VDF_STACK_VIDPROC, no GPU work */

/* ------------------------------------------------------------------------------------ context */

/* Binds a context to CUDA device `device_id` (one process per GPU). */
int vdf_ctx_create(int device_id, vdf_ctx** out);
/* One context over n_dev (1..8) distinct GPUs of one node, driven from this process: peer access is
 * enabled between them and vdf_search, vdf_search_with_references and vdf_hash_stacks use all of them transparently -- the
 * caller's table is staged and sorted once on dev_ids[0] and copied GPU -> GPU, every device evaluates its share of the
 * pair matrix (one host thread per device for the duration of the call), matches travel through the fused peer-memory
 * exchange, grouping runs on dev_ids[0]; stacks are hashed in contiguous shards.  Results are identical to a one-device
 * context's.  All other entry points act on dev_ids[0].  vdf_ctx_set_option reaches every device. */
int vdf_ctx_create_multi(const int* dev_ids, int n_dev, vdf_ctx** out);
int vdf_ctx_device_count(const vdf_ctx* ctx);
void vdf_ctx_destroy(vdf_ctx* ctx);
const char* vdf_last_error(const vdf_ctx* ctx);

/* Multi-GPU: this context evaluates only its share (rank of world) of the pair-matrix tile blocks of
 * vdf_search_self* / the candidate tiles of vdf_search_refs*; edge lists of all ranks are concatenated by
 * the caller (NCCL all-gather) before vdf_group_greedy*.  Default rank 0 of 1. */
int vdf_ctx_set_shard(vdf_ctx* ctx, uint32_t rank, uint32_t world);

/* Edge exchange of a multi-GPU search over NVLink peer memory, fused into the pair kernel (search_variant 6): instead of
 * gathering per-rank edge lists with a collective, the kernel that finds a match takes a slot
 * from its local counter and stores the key into its own segment of EVERY rank's exchange buffer (plain remote stores over
 * NVLink, overlapped with the math); a flag barrier in peer memory publishes the per-rank counts and ends the call, after
 * which vdf_search_self_device / vdf_search_refs_device return the sorted keys of ALL ranks on every rank.  One process per
 * GPU, <= 8 GPUs of one node.
 *   vdf_peer_alloc: (re)allocates this rank's buffer for capacity_keys matches (total over all ranks; each rank may
 *                   contribute capacity_keys / world) and writes its CUDA IPC handle (64 bytes);
 *   the caller exchanges the handles of all ranks (any control-plane channel; dist.py uses all_gather_object);
 *   vdf_peer_open:  maps the peers' buffers (handles = world x 64 bytes in rank order) - every rank must have returned from
 *                   vdf_peer_open before any rank searches (a host barrier);
 *   option "exchange" = 1 switches the searches to the exchange; all ranks must then make the same sequence of search calls
 *                   (how the work is divided - vdf_ctx_set_shard for the self search, candidate slices for the reference
 *                   search - is independent of it).  VDF_ERR_EDGE_OVERFLOW is returned on every rank alike, with the same
 *                   count, when some rank's matches exceed its segment: re-allocate larger on all ranks and repeat. */
int vdf_peer_alloc(vdf_ctx* ctx, uint64_t capacity_keys, unsigned char handle_out[64]);
int vdf_peer_open(vdf_ctx* ctx, uint32_t rank, uint32_t world, const unsigned char* handles);
int vdf_peer_close(vdf_ctx* ctx);

/* Tuning knobs: "max_edges" (edge-buffer growth cap, default 2^28, at most 2^32 - 2), "initial_edges" (default 2^22),
 * "search_variant": 0 = XOR+POPC; 1, 2 = XOR + carry-save adders + POPC (8x8 / 8x4 pairs per thread);
 *   5 = tcgen05.mma kind::i8 on CTA pairs (cta_group::2), packed tiles, bits expanded to bytes inside the kernel;
 *   6 (default) = the same with the bits expanded to e2m1 nibbles and tcgen05.mma kind::mxf4 (twice the kind::i8 rate).
 *   All five are bit-identical.  (3 and 4 of round 1, byte-expanded tiles in HBM, were removed: strictly dominated.)
 * "tc_chunk": column super-tiles per work unit of variants 5, 6 (0 = automatic); "tc_unit_order": variant-6 work-unit order
 * (0 chunk-major, 1 row-pair-major); "tc_a_tmem": variant 6 keeps three quarters of the row operand in tensor memory (1,
 * default) or all of it in shared memory (0); "tc_fold": -1 (default) variant 6 folds the column popcounts into the free K
 * positions 1000..1023 whenever no hash of either operand sets those bits (every real VideoHash: dct_3d.rs:55-66), which
 * makes its 64-column screen exact at any tolerance; 0 = always the popcount-screen epilogue;
 * "peer_timeout_ms": how long the peer exchange waits for the slowest rank (0 = automatic, grows with the problem);
 * "hash_fused": 1 (default) a hashing call is ONE persistent kernel (letterbox scan, crop, resize, DCT, pack: hash_fused_kernel),
 * 0 = the per-frame kernels it replaced (letterbox scan kernels, then one resize block per frame), for which "hash_variant"
 * picks the resize kernel and "hash_chunks" (1..4, default 1) / "hash_overlap" (default 0) split the call into chunks of
 * stacks with the scan of chunk k+1 beside the resize of chunk k on a second stream (measured slower; experiment knobs);
 * "hash_fuse_dct": 1 (default) DCT + pack inside the hashing kernel, 0 = a kernel of its own. */
int vdf_ctx_set_option(vdf_ctx* ctx, const char* key, int64_t value);

/* The cudaStream_t all kernels of this context are launched on (for CUDA-event timing by the caller). */
void* vdf_ctx_stream(vdf_ctx* ctx);

/* Counters since context creation: kernels launched by this library, and bytes copied H2D / D2H. */
void vdf_ctx_counters(const vdf_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* Device time (CUDA events on the context's stream) accumulated over the launches of one of the dominant
 * kernels: which = 0 hamming tiles, 1 crop+resize, 2 letterbox scan, 3 DCT+pack.  reset != 0 clears the slot. */
int vdf_ctx_kernel_time(vdf_ctx* ctx, int which, double* total_ms, uint64_t* launches, int reset);

/* ------------------------------------------------------------------------------------ search path */

/* Replaces the comparison loop of Search::search_self (search_algorithm.rs:81-117,140-156) together with
 * VideoHash::hamming_distance (video_hash.rs:190-192,311-317): every pair i<j with
 * dur[j] <= (f64(dur[i]) * 1.1) as u32 and hamming_1024(i,j) <= tol_int becomes an edge.
 * tol_int = (tolerance * 1000.0) as u32 is computed by the caller (search_algorithm.rs:82). */
int vdf_search_self(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n,
                    uint32_t tol_int, vdf_edges* out);

/* Replaces the consumption rule of Search::search_self (search_algorithm.rs:131-170): targets are taken in
 * ascending order, each consumes its still-unmatched neighbours; output order as the reference's
 * Vec<Vec<PathBuf>> after ret.reverse().  Runs on the GPU (parallel rounds of the greedy rule). */
int vdf_group_greedy(vdf_ctx* ctx, uint64_t n, const vdf_edges* edges, vdf_groups* out);

/* OPTIONAL, not reference behaviour: connected components of the edge graph by a lock-free GPU union-find -- what the
 * app's DisjointSet (vid_dup_finder_app/src/app/disjoint_set.rs:22-44) computes for confirmed pairs, and a superset of
 * every greedy group.  Same output convention as vdf_group_greedy (members ascending, the component's smallest entry
 * last, groups by descending smallest entry).  With the context option "grouping" = 1, vdf_search and
 * vdf_search_self_groups group this way instead of by the reference's rule (default 0). */
int vdf_group_components(vdf_ctx* ctx, uint64_t n, const vdf_edges* edges, vdf_groups* out);
int vdf_group_components_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys, uint64_t n_edges, vdf_groups* out);

/* vdf_search_self + vdf_group_greedy with the edge list kept in HBM: what `search()` (video_dup_finder.rs:7-13)
 * calls.  Groups of fewer than 2 entries cannot occur (MatchGroup::new, match_group.rs:21-30). */
int vdf_search_self_groups(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n,
                           uint32_t tol_int, vdf_groups* out);

/* Replaces search_with_references' inner loop (video_dup_finder.rs:19-46; Search::search_one and
 * duration_slice, search_algorithm.rs:63-77,173-185; consume = false): for each reference r, all candidates k with
 * (f64(ref_dur[r])*0.95) as u32 <= cand_dur[k] <= (f64(ref_dur[r])*1.05) as u32 and hamming <= tol_int.
 * Rows with no match are empty (the caller skips them, video_dup_finder.rs:38-43). */
int vdf_search_refs(vdf_ctx* ctx, const uint64_t* cand_sorted, const uint32_t* cand_dur_sorted, uint64_t n_cand,
                    const uint64_t* refs, const uint32_t* ref_dur, uint64_t n_ref, uint32_t tol_int, vdf_csr* out);

/* Device-resident forms (HBM in, HBM out) used by pipelines that keep the table on the GPU and by the
 * multi-GPU plumbing.  Edges are returned as sorted u64 keys (i << 32 | j), resp. (ref << 32 | cand).
 * `capacity` is the size of d_keys_out in keys; if more matches exist, *n_out receives the required count
 * and the call returns VDF_ERR_EDGE_OVERFLOW.  cand_index_base is added to candidate indices (a rank that
 * holds a contiguous slice of the sorted corpus passes the slice's first global index). */
int vdf_search_self_device(vdf_ctx* ctx, const uint64_t* d_hash_sorted, const uint32_t* d_dur_sorted, uint64_t n,
                           uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int vdf_search_refs_device(vdf_ctx* ctx, const uint64_t* d_cand_sorted, const uint32_t* d_cand_dur_sorted,
                           uint64_t n_cand, uint64_t cand_index_base, const uint64_t* d_refs,
                           const uint32_t* d_ref_dur, uint64_t n_ref, uint32_t tol_int, uint64_t* d_keys_out,
                           uint64_t capacity, uint64_t* n_out);
/* d_remap (optional, n x u32 in HBM): every index written to `out` is d_remap[sorted position] -- how vdf_search hands back
 * the caller's indices without a host pass.  Groups by the context's "grouping" option (default: the reference's rule). */
int vdf_group_greedy_device(vdf_ctx* ctx, uint64_t n, const uint64_t* d_keys_sorted, uint64_t n_edges, const uint32_t* d_remap,
                            vdf_groups* out);

/* ---- prepared tables ----------------------------------------------------------------------------------------------
 * `Search::from(hashes)` (search_algorithm.rs:188-198) builds the sorted table once; `search_self(tolerance)` and
 * `search_with_references` (:81, :40) then run on it any number of times.  A vdf_table is that object with the table
 * resident in HBM in the pair kernels' layout (packed tiles, popcounts, fold units) together with the duration windows and
 * the work-unit list of the context's current shard: searching it launches the pair kernel, reads one match count and
 * groups -- nothing is re-packed.  vdf_table_create takes host arrays in sorted order and owns its copy;
 * vdf_table_create_device takes device arrays that must stay valid and unchanged while the table lives.  A table belongs
 * to the context it was created with; destroy it before the context. */
int vdf_table_create(vdf_ctx* ctx, const uint64_t* hash_sorted, const uint32_t* dur_sorted, uint64_t n, vdf_table** out);
int vdf_table_create_device(vdf_ctx* ctx, const uint64_t* d_hash_sorted, const uint32_t* d_dur_sorted, uint64_t n, vdf_table** out);
void vdf_table_destroy(vdf_table* table);
uint64_t vdf_table_len(const vdf_table* table);
/* as vdf_search_self_device / vdf_search_self_groups / vdf_search_refs_device, on a prepared table */
int vdf_table_search_self_device(vdf_table* table, uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);
int vdf_table_search_self_groups(vdf_table* table, uint32_t tol_int, vdf_groups* out);
int vdf_table_search_refs_device(vdf_table* cand_table, uint64_t cand_index_base, const uint64_t* d_refs, const uint32_t* d_ref_dur,
                                 uint64_t n_ref, uint32_t tol_int, uint64_t* d_keys_out, uint64_t capacity, uint64_t* n_out);

/* Number of (i,j) pairs inside the duration windows, i.e. how many hamming_distance calls the reference's
 * search_self would make with nothing consumed (throughput accounting). */
int vdf_self_window_pairs(vdf_ctx* ctx, const uint32_t* dur_sorted, uint64_t n, uint64_t* pairs_out);

/* ---- the crate's two public search functions, whole ------------------------------------------------------------
 * Inputs in the CALLER's order, struct-of-arrays: hashes [n][16] u64, durations [n] u32, and the src_paths as one
 * byte blob with n+1 offsets (path i = path_blob[path_off[i] .. path_off[i+1]), no terminators).  The library does
 * what video_dup_finder.rs does around the comparison loops: the stable (duration, src_path) sort of Search::sort
 * (search_algorithm.rs:55-61; Rust Unix `Path` ordering = component-wise), tol_int = (tolerance * 1000.0) as u32
 * (search_algorithm.rs:82), the GPU search, and the mapping back, so every index that comes out refers to the
 * caller's arrays.  Host threads only cut 20-byte sort keys and copy the hashes to pinned memory; the sort itself, the
 * gather into sorted order and the mapping back run on the GPU (csrc/host.cu). */

/* Replaces Search::sort (search_algorithm.rs:55-61): order_out[k] = index of the k-th entry in sorted order. */
int vdf_sort_order(const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n,
                   uint64_t* order_out);

/* Search::sort for a table whose hashes already live in HBM (e.g. straight out of vdf_hash_stacks_device): the permutation
 * (n x u32) and, optionally, the durations in sorted order are written to DEVICE memory; the caller gathers its hash rows by
 * the permutation (or hands it to a kernel of its own).  Sort keys are cut on the host, the sort runs on the GPU. */
int vdf_sort_order_device(vdf_ctx* ctx, const uint32_t* durations, const char* path_blob, const uint64_t* path_off, uint64_t n,
                          uint32_t* d_order_out, uint32_t* d_dur_sorted_out);

/* `Search::seed` + `sort` (search_algorithm.rs:31-34,55-61) with the sorted table left RESIDENT in HBM: the stable
 * (duration, Path) permutation goes to order_out[n]; hashes and durations are gathered in that order through pinned
 * memory and uploaded on the context's stream - into d_hash_dst / d_dur_dst (n x 128 B / n x 4 B of device memory owned by
 * the caller) when given, else into buffers owned by the context, valid until its next vdf_stage_sorted / vdf_search* call.
 * *d_hash_sorted / *d_dur_sorted receive the device pointers; they feed vdf_search_self_device.  In a multi-GPU search one
 * rank stages and broadcasts the table, the others skip the host sort (vid_dup_finder_lib_b200/dist.py). */
int vdf_stage_sorted(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob,
                     const uint64_t* path_off, uint64_t n, uint64_t* order_out, uint64_t* d_hash_dst, uint32_t* d_dur_dst,
                     const uint64_t** d_hash_sorted, const uint32_t** d_dur_sorted);

/* Replaces `search(hashes, tolerance)` (video_dup_finder.rs:7-13): groups exactly as the reference returns them
 * (matches in sorted order, the target last, groups by descending target; every group has >= 2 entries), with
 * member_idx holding the caller's indices.  The caller builds MatchGroup::new(paths of members). */
int vdf_search(vdf_ctx* ctx, const uint64_t* hashes, const uint32_t* durations, const char* path_blob,
               const uint64_t* path_off, uint64_t n, double tolerance, vdf_groups* out);

/* Replaces `search_with_references(ref_hashes, new_hashes, tolerance)` (video_dup_finder.rs:19-46): row r = the
 * entries of new_hashes (caller's indices, in sorted order) matching reference r inside its duration slice; empty
 * rows are the references the caller skips (video_dup_finder.rs:38-43). */
int vdf_search_with_references(vdf_ctx* ctx, const uint64_t* ref_hashes, const uint32_t* ref_durations, uint64_t n_ref,
                               const uint64_t* new_hashes, const uint32_t* new_durations, const char* new_path_blob,
                               const uint64_t* new_path_off, uint64_t n_new, double tolerance, vdf_csr* out);

/* Wall-clock milliseconds of the phases of the last vdf_search / vdf_search_with_references / vdf_stage_sorted call:
 * [0] sort keys cut on the host, [1] hashes to pinned memory + uploads + GPU sort, [2] device work incl. result D2H,
 * [3] 0 (the index remap is part of the device work since round 2). */
int vdf_ctx_last_phases(const vdf_ctx* ctx, double* ms4);

/* ---- the application's hash cache file (host code, csrc/cache.cu) ----------------------------------------------
 * The app stores HashMap<PathBuf, MtimeCacheEntry<Result<VideoHash, Error>>> with bincode 2 `standard()`
 * (vid_dup_finder_app/src/video_hash_filesystem_cache/generic_filesystem_cache/base_fs_cache.rs:26,106-112,192-196;
 * processing_fs_cache.rs:23-27; generic_cache_if.rs:22-23; video_hash.rs:26-32; video_hashing/mod.rs:17-28) and loads it
 * whole before every search.  vdf_cache_load reads such a file into struct-of-arrays, entry i in file order:
 *   kind[i]                     VDF_CACHE_OK or which Error the entry caches
 *   hashes[i][16], durations[i] the VideoHash (zero for error entries)
 *   key_*                       the map key (the file's path); src_*: VideoHash.src_path (equal to the key when the app
 *                               wrote the file, kept separately so that a load/save round trip is byte-faithful)
 *   msg_*                       the String of Error::VidProc
 *   mtime_secs/nanos[i]         MtimeCacheEntry.cache_mtime as serde writes a SystemTime
 * so hashes/durations/src_* of the VDF_CACHE_OK entries go to vdf_search without a per-entry object ever existing.
 * vdf_cache_save writes the same encoding (temporary file + rename, like base_fs_cache.rs:84,157). */
#define VDF_CACHE_OK 0
#define VDF_CACHE_ERR_NOT_VIDEO 1         /* Error::NotVideo */
#define VDF_CACHE_ERR_VIDPROC 2           /* Error::VidProc(String) */
#define VDF_CACHE_ERR_NOT_ENOUGH_FRAMES 3 /* Error::NotEnoughFrames */
typedef struct {
    uint64_t n;
    int32_t* kind;
    uint64_t* hashes;
    uint32_t* durations;
    char* key_blob;
    uint64_t* key_off; /* n + 1 */
    char* src_blob;
    uint64_t* src_off; /* n + 1 */
    char* msg_blob;
    uint64_t* msg_off; /* n + 1 */
    uint64_t* mtime_secs;
    uint32_t* mtime_nanos;
} vdf_cache;
int vdf_cache_load(const char* file, vdf_cache* out);
int vdf_cache_save(const char* file, const vdf_cache* cache);
void vdf_free_cache(vdf_cache* cache);

void vdf_free_edges(vdf_edges* e);
void vdf_free_groups(vdf_groups* g);
void vdf_free_csr(vdf_csr* c);

/* ------------------------------------------------------------------------------------ hashing path */

/* Replaces the compute tail of gen_hash (video_hash_builder.rs:214-223): crop_video_frames (:188-212) with
 * cropdetect_letterbox (vid_dup_finder_common/src/video_frames_gray.rs:38-128,201-210; Crop::union
 * crop.rs:53-68), crop_resize_buf to 16x16 (vid_dup_finder_common/src/resize_gray.rs:11-54, fast_image_resize
 * Lanczos3 u8), Dct3d::from_images + dct_3d (dct_3d.rs:15-53, raw_dct_ops.rs:107-142, f64) and the
 * threshold / bit pack (dct_3d.rs:55-66, video_hash.rs:63-70).
 * frames: host pointer; stacks are staged to HBM with pinned async copies.  out_hash: n x 16 u64.
 * out_status: n x i32 (VDF_STACK_*).  out_crop: optional n x 4 u32 (left, right, top, bottom).
 * cropdetect: VDF_CROPDETECT_NONE / _LETTERBOX / _MOTION. */
int vdf_hash_stacks(vdf_ctx* ctx, const uint8_t* frames, const vdf_stack_desc* desc, uint32_t n, int cropdetect,
                    uint64_t* out_hash, int32_t* out_status, uint32_t* out_crop);

/* Same with frames already in HBM; d_out_hash is a device pointer (n x 16 u64), status/crop are host. */
int vdf_hash_stacks_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n,
                           int cropdetect, uint64_t* d_out_hash, int32_t* out_status, uint32_t* out_crop);

/* Debug / parity taps: the 16x16x16 u8 cube after crop+resize ([t][row][col]) for each stack. */
int vdf_hash_stacks_small_device(vdf_ctx* ctx, const uint8_t* d_frames, const vdf_stack_desc* desc, uint32_t n,
                                 int cropdetect, uint8_t* d_out_small /* n x 4096 */, uint32_t* out_crop);

/* Hash from an already-resized cube (Dct3d::from_images + hash_bits only). */
int vdf_hash_from_small(vdf_ctx* ctx, const uint8_t* small /* host, n x 4096 */, uint32_t n, uint64_t* out_hash);

/* ---- batch-oriented hashing (csrc/pipeline.cu) ------------------------------------------------------------------
 * Replaces the app's per-file loop (video_hash_filesystem_cache.rs:237-257: one rayon worker decodes AND hashes one
 * file): decode threads push their <= 16 decoded gray frames, a worker thread owned by the pipeline hashes them in
 * batches through vdf_hash_stacks, a collector polls results.  vdf_pipeline_push may be called from any number of
 * threads and copies the frames into pinned memory before it returns (the caller may reuse its buffers); it blocks
 * while both batch buffers are full.  While the pipeline exists the context must not be used by other threads.
 * status/hash/crop of a result are exactly vdf_hash_stacks' outputs for that stack; a negative status is a library
 * error (vdf_pipeline_error). */
typedef struct vdf_hash_pipeline vdf_hash_pipeline;
typedef struct {
    uint64_t tag;      /* the caller's id of the video */
    int32_t status;    /* VDF_STACK_OK / VDF_STACK_NOT_ENOUGH_FRAMES / VDF_STACK_VIDPROC, or a negative VDF_ERR_* */
    uint32_t crop[4];  /* left, right, top, bottom */
    uint64_t hash[16]; /* VideoHash.hash */
} vdf_pipeline_result;
int vdf_pipeline_create(vdf_ctx* ctx, uint32_t max_batch_stacks, uint64_t batch_bytes, int cropdetect, vdf_hash_pipeline** out);
int vdf_pipeline_push(vdf_hash_pipeline* p, uint64_t tag, const uint8_t* const* frames, uint32_t n_frames, uint32_t width,
                      uint32_t height, uint32_t pitch, uint32_t flags);
/* submit the partly filled batch and wait until everything pushed so far has a result */
int vdf_pipeline_flush(vdf_hash_pipeline* p);
/* take up to max_results finished results; wait != 0 blocks until there is one (or nothing is in flight) */
int vdf_pipeline_poll(vdf_hash_pipeline* p, vdf_pipeline_result* out, uint32_t max_results, uint32_t* n_out, int wait);
const char* vdf_pipeline_error(const vdf_hash_pipeline* p);
void vdf_pipeline_destroy(vdf_hash_pipeline* p);

const char* vdf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VDF_B200_H */
