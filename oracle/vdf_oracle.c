/*
 * vdf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See vdf_oracle.h.
 *
 * Restates, function by function, the reference's hot paths.  Citations are file:line relative to
 * /root/reference.  Compile with -ffp-contract=off: the reference (rustc) never fuses a*b+c.
 */
#include "vdf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== search path == */

/* video_hash.rs:311-317 */
uint32_t vdfo_hamming(const uint64_t* x, const uint64_t* y) {
    uint32_t acc = 0;
    for (int w = 0; w < VDFO_HASH_WORDS; ++w) acc += (uint32_t)__builtin_popcountll(x[w] ^ y[w]);
    return acc;
}

/* Rust `f64 as u32`: truncate toward zero, saturate, NaN -> 0 */
static uint32_t f64_as_u32(double v) {
    if (!(v == v)) return 0u;
    if (v <= 0.0) return 0u;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

/* search_algorithm.rs:82 with definitions.rs:40 (TOLERANCE_SCALING_FACTOR = 1000.0) */
uint32_t vdfo_tolerance_int(double tolerance) { return f64_as_u32(tolerance * 1000.0); }

/* search_algorithm.rs:99 */
uint32_t vdfo_self_window_thresh(uint32_t duration) { return f64_as_u32((double)duration * 1.1); }

/* search_algorithm.rs:174,179 */
void vdfo_ref_window_durations(uint32_t duration, uint32_t* lo, uint32_t* hi) {
    *lo = f64_as_u32((double)duration * 0.95);
    *hi = f64_as_u32((double)duration * 1.05);
}

/* ---- Rust Path ordering (Unix): compare component lists; Component variant order is
 * RootDir < CurDir < ParentDir < Normal(bytes).  "." is dropped except as the leading component of a
 * relative path; empty pieces (repeated or trailing '/') are dropped. */
typedef struct {
    const char* p;
    size_t pos, len;
    int started, has_root;
} comp_iter;
typedef struct {
    int kind; /* 1 RootDir, 2 CurDir, 3 ParentDir, 4 Normal */
    const char* s;
    size_t n;
} comp;

static void comp_iter_init(comp_iter* it, const char* p) {
    it->p = p;
    it->len = strlen(p);
    it->pos = 0;
    it->started = 0;
    it->has_root = (it->len > 0 && p[0] == '/');
}

static int comp_next(comp_iter* it, comp* out) {
    if (!it->started) {
        it->started = 1;
        if (it->has_root) {
            out->kind = 1;
            out->s = it->p;
            out->n = 1;
            it->pos = 1;
            return 1;
        }
        /* leading "." of a relative path is kept as CurDir */
        if (it->len >= 1 && it->p[0] == '.' && (it->len == 1 || it->p[1] == '/')) {
            out->kind = 2;
            out->s = it->p;
            out->n = 1;
            it->pos = 1;
            return 1;
        }
    }
    for (;;) {
        while (it->pos < it->len && it->p[it->pos] == '/') it->pos++;
        if (it->pos >= it->len) return 0;
        size_t s = it->pos;
        while (it->pos < it->len && it->p[it->pos] != '/') it->pos++;
        size_t n = it->pos - s;
        if (n == 1 && it->p[s] == '.') continue;
        out->s = it->p + s;
        out->n = n;
        out->kind = (n == 2 && it->p[s] == '.' && it->p[s + 1] == '.') ? 3 : 4;
        return 1;
    }
}

int vdfo_path_cmp(const char* a, const char* b) {
    comp_iter ia, ib;
    comp ca, cb;
    comp_iter_init(&ia, a);
    comp_iter_init(&ib, b);
    for (;;) {
        int ha = comp_next(&ia, &ca), hb = comp_next(&ib, &cb);
        if (!ha && !hb) return 0;
        if (!ha) return -1;
        if (!hb) return 1;
        if (ca.kind != cb.kind) return ca.kind < cb.kind ? -1 : 1;
        if (ca.kind == 4) {
            size_t m = ca.n < cb.n ? ca.n : cb.n;
            int c = memcmp(ca.s, cb.s, m);
            if (c) return c < 0 ? -1 : 1;
            if (ca.n != cb.n) return ca.n < cb.n ? -1 : 1;
        }
    }
}

/* search_algorithm.rs:55-61 : sort_by_key is a stable merge sort on (duration, path) */
typedef struct {
    const uint32_t* dur;
    const char* const* paths;
} sort_ctx;

static int key_le(const sort_ctx* c, uint64_t a, uint64_t b) { /* key(a) <= key(b) */
    if (c->dur[a] != c->dur[b]) return c->dur[a] < c->dur[b];
    return vdfo_path_cmp(c->paths[a], c->paths[b]) <= 0;
}

static void msort(const sort_ctx* c, uint64_t* v, uint64_t* tmp, uint64_t n) {
    if (n < 2) return;
    uint64_t h = n / 2;
    msort(c, v, tmp, h);
    msort(c, v + h, tmp, n - h);
    uint64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = key_le(c, v[i], v[j]) ? v[i++] : v[j++];
    while (i < h) tmp[k++] = v[i++];
    while (j < n) tmp[k++] = v[j++];
    memcpy(v, tmp, n * sizeof(uint64_t));
}

void vdfo_sort_order(const uint32_t* duration, const char* const* paths, uint64_t n, uint64_t* order_out) {
    sort_ctx c = {duration, paths};
    for (uint64_t i = 0; i < n; ++i) order_out[i] = i;
    uint64_t* tmp = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    msort(&c, order_out, tmp, n);
    free(tmp);
}

/* growable u64 vector */
typedef struct {
    uint64_t* d;
    uint64_t n, cap;
} vec64;
static int vpush(vec64* v, uint64_t x) {
    if (v->n == v->cap) {
        uint64_t nc = v->cap ? v->cap * 2 : 64;
        uint64_t* nd = (uint64_t*)realloc(v->d, nc * sizeof(uint64_t));
        if (!nd) return -1;
        v->d = nd;
        v->cap = nc;
    }
    v->d[v->n++] = x;
    return 0;
}

/* reverse the order of CSR groups in place (search_algorithm.rs:136,167 ret.reverse()) */
static int finish_groups(vec64* ptr, vec64* mem, uint64_t** gp_out, uint64_t** m_out, uint64_t* ng_out) {
    uint64_t ng = ptr->n ? ptr->n - 1 : 0;
    uint64_t* gp = (uint64_t*)malloc((ng + 1) * sizeof(uint64_t));
    uint64_t* mm = (uint64_t*)malloc((mem->n ? mem->n : 1) * sizeof(uint64_t));
    if (!gp || !mm) return -1;
    uint64_t o = 0;
    gp[0] = 0;
    for (uint64_t g = 0; g < ng; ++g) {
        uint64_t src = ng - 1 - g;
        uint64_t a = ptr->d[src], b = ptr->d[src + 1];
        memcpy(mm + o, mem->d + a, (b - a) * sizeof(uint64_t));
        o += b - a;
        gp[g + 1] = o;
    }
    free(ptr->d);
    free(mem->d);
    *gp_out = gp;
    *m_out = mm;
    *ng_out = ng;
    return 0;
}

/* search_algorithm.rs:81-171, restated with explicit cursors and a matched[] array */
int vdfo_search_self(const uint64_t* H, const uint32_t* dur, uint64_t n, uint32_t tol_int, uint64_t** gp_out,
                     uint64_t** m_out, uint64_t* ng_out) {
    vec64 ptr = {0, 0, 0}, mem = {0, 0, 0};
    if (vpush(&ptr, 0)) return -1;
    if (n == 0) return finish_groups(&ptr, &mem, gp_out, m_out, ng_out); /* :88-90 */
    uint8_t* matched = (uint8_t*)calloc(n, 1);
    if (!matched) return -1;
    uint64_t lhs = 0, rhs = 0;
    for (;;) {
        /* advance_rhs :93-117 : skip matched entries; stop at the first unmatched entry whose duration
         * exceeds the threshold, or at the end of the vector */
        uint32_t thresh = vdfo_self_window_thresh(dur[lhs]);
        while (rhs < n) {
            if (!matched[rhs] && dur[rhs] > thresh) break;
            rhs++;
        }
        if (lhs < rhs) { /* :140 */
            matched[lhs] = 1; /* :147 */
            uint64_t start = mem.n;
            for (uint64_t c = lhs + 1; c < rhs; ++c) { /* :150-156 */
                if (!matched[c] && vdfo_hamming(H + 16 * lhs, H + 16 * c) <= tol_int) {
                    if (vpush(&mem, c)) return -1;
                    matched[c] = 1;
                }
            }
            if (mem.n != start) { /* :158-161 target goes last */
                if (vpush(&mem, lhs)) return -1;
                if (vpush(&ptr, mem.n)) return -1;
            }
        }
        /* advance_lhs :119-129 : next unmatched entry, or finish */
        do {
            lhs++;
        } while (lhs < n && matched[lhs]);
        if (lhs >= n) break;
    }
    free(matched);
    return finish_groups(&ptr, &mem, gp_out, m_out, ng_out);
}

/* SURVEY A.2 edge predicate, brute force, (i,j)-sorted by construction */
int vdfo_self_edges(const uint64_t* H, const uint32_t* dur, uint64_t n, uint32_t tol_int, uint64_t** e_out,
                    uint64_t* ne_out) {
    vec64 e = {0, 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t thresh = vdfo_self_window_thresh(dur[i]);
        for (uint64_t j = i + 1; j < n && dur[j] <= thresh; ++j) {
            if (vdfo_hamming(H + 16 * i, H + 16 * j) <= tol_int) {
                if (vpush(&e, i) || vpush(&e, j)) return -1;
            }
        }
    }
    *e_out = e.d ? e.d : (uint64_t*)malloc(8);
    *ne_out = e.n / 2;
    return 0;
}

uint64_t vdfo_self_window_pairs(const uint32_t* dur, uint64_t n) {
    uint64_t pairs = 0, j = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t thresh = vdfo_self_window_thresh(dur[i]);
        if (j < i + 1) j = i + 1;
        while (j < n && dur[j] <= thresh) j++;
        pairs += j - (i + 1);
    }
    return pairs;
}

/* greedy rule over sorted edges: i ascending; an unconsumed i becomes a target and consumes its
 * unconsumed neighbours j>i in ascending j (search_algorithm.rs:147-161) */
int vdfo_group_from_edges(uint64_t n, const uint64_t* edges, uint64_t ne, uint64_t** gp_out, uint64_t** m_out,
                          uint64_t* ng_out) {
    vec64 ptr = {0, 0, 0}, mem = {0, 0, 0};
    if (vpush(&ptr, 0)) return -1;
    uint8_t* consumed = (uint8_t*)calloc(n ? n : 1, 1);
    if (!consumed) return -1;
    uint64_t k = 0;
    while (k < ne) {
        uint64_t i = edges[2 * k];
        uint64_t k_end = k;
        while (k_end < ne && edges[2 * k_end] == i) k_end++;
        if (!consumed[i]) {
            consumed[i] = 1;
            uint64_t start = mem.n;
            for (uint64_t q = k; q < k_end; ++q) {
                uint64_t j = edges[2 * q + 1];
                if (!consumed[j]) {
                    consumed[j] = 1;
                    if (vpush(&mem, j)) return -1;
                }
            }
            if (mem.n != start) {
                if (vpush(&mem, i) || vpush(&ptr, mem.n)) return -1;
            }
        }
        k = k_end;
    }
    free(consumed);
    return finish_groups(&ptr, &mem, gp_out, m_out, ng_out);
}

/* slice::partition_point over the sorted durations */
static uint64_t pp_lt(const uint32_t* d, uint64_t n, uint32_t v) { /* first k with !(d[k] < v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (d[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
static uint64_t pp_le(const uint32_t* d, uint64_t n, uint32_t v) { /* first k with !(d[k] <= v) */
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (d[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

/* video_dup_finder.rs:19-46 ; search_algorithm.rs:63-77 (consume=false), :173-185 */
int vdfo_search_refs(const uint64_t* C, const uint32_t* cdur, uint64_t nc, const uint64_t* R, const uint32_t* rdur,
                     uint64_t nr, uint32_t tol_int, uint64_t** rp_out, uint64_t** ci_out) {
    uint64_t* rp = (uint64_t*)malloc((nr + 1) * sizeof(uint64_t));
    if (!rp) return -1;
    vec64 ci = {0, 0, 0};
    rp[0] = 0;
    for (uint64_t r = 0; r < nr; ++r) {
        uint32_t lo_d, hi_d;
        vdfo_ref_window_durations(rdur[r], &lo_d, &hi_d);
        uint64_t lo = pp_lt(cdur, nc, lo_d), hi = pp_le(cdur, nc, hi_d);
        for (uint64_t k = lo; k < hi; ++k)
            if (vdfo_hamming(R + 16 * r, C + 16 * k) <= tol_int)
                if (vpush(&ci, k)) return -1;
        rp[r + 1] = ci.n;
    }
    *rp_out = rp;
    *ci_out = ci.d ? ci.d : (uint64_t*)malloc(8);
    return 0;
}

void vdfo_free(void* p) { free(p); }

/* ======================================================================= hashing path == */

/* video_frames_gray.rs:66-103 : is this 1-px strip letterbox?  strip = len pixels at `step` stride */
static int strip_is_letterbox(const uint8_t* p, size_t step, uint32_t len, int mode, uint8_t tol) {
    size_t matching = 0;
    if (mode == VDFO_LB_BLACKWHITE) { /* :71-79 */
        for (uint32_t k = 0; k < len; ++k) {
            uint8_t l = p[k * step];
            if (l <= tol || l >= (uint8_t)(255 - tol)) matching++;
        }
    } else { /* :80-97 */
        size_t hist[256];
        memset(hist, 0, sizeof hist);
        for (uint32_t k = 0; k < len; ++k) hist[p[k * step]]++;
        int mode_v = 0; /* Iterator::max_by_key keeps the LAST maximum */
        for (int v = 0; v < 256; ++v)
            if (hist[v] >= hist[mode_v]) mode_v = v;
        for (uint32_t k = 0; k < len; ++k) {
            int d = (int)p[k * step] - mode_v;
            if (d < 0) d = -d;
            if (d <= (int)tol) matching++;
        }
    }
    double proportion = (double)matching / (double)len; /* :99-100 */
    return proportion > 0.9;
}

void vdfo_letterbox_frame(const uint8_t* pix, uint32_t w, uint32_t h, size_t pitch, int mode, uint8_t tol,
                          uint32_t out[4]) {
    uint32_t l = 0, r = 0, t = 0, b = 0;
    /* :105-111 take_while(is_letterbox).count() from each edge inwards, full-length strips */
    while (l < w && strip_is_letterbox(pix + l, pitch, h, mode, tol)) l++;
    while (r < w && strip_is_letterbox(pix + (w - r - 1), pitch, h, mode, tol)) r++;
    while (t < h && strip_is_letterbox(pix + (size_t)t * pitch, 1, w, mode, tol)) t++;
    while (b < h && strip_is_letterbox(pix + (size_t)(h - b - 1) * pitch, 1, w, mode, tol)) b++;
    int rem_h = (int)w - (int)l - (int)r, rem_v = (int)h - (int)t - (int)b; /* :119-127 */
    if (rem_h >= 1 && rem_v >= 1) {
        out[0] = l, out[1] = r, out[2] = t, out[3] = b;
    } else {
        out[0] = out[1] = out[2] = out[3] = 0;
    }
}

int vdfo_cropdetect_letterbox(const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h, size_t pitch,
                              size_t frame_stride, uint32_t out[4]) {
    if (n_frames == 0) return VDFO_NOT_ENOUGH_FRAMES; /* reduce() on an empty iterator -> None */
    int first = 1;
    uint32_t taken = 0;
    for (uint32_t f = 0; f < n_frames && taken < 8; f += 8, ++taken) { /* step_by(8).take(8) :204 */
        uint32_t c[4];
        vdfo_letterbox_frame(frames + (size_t)f * frame_stride, w, h, pitch, VDFO_LB_ANYCOLOUR, 16, c);
        if (first) {
            memcpy(out, c, sizeof c);
            first = 0;
        } else { /* Crop::union crop.rs:53-68 = per-side min */
            for (int s = 0; s < 4; ++s)
                if (c[s] < out[s]) out[s] = c[s];
        }
    }
    return VDFO_OK;
}

/* ---- fast_image_resize 5.1, U8, ResizeAlg::Convolution(FilterType::Lanczos3) ---------------------- */

static double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x *= 3.14159265358979323846;
    return sin(x) / x;
}
static double lanczos3_filter(double x) { /* truncated sinc, support 3 */
    if (x >= -3.0 && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3.0);
    return 0.0;
}

typedef struct {
    uint32_t window, precision;
    uint32_t* bounds; /* out_size x (start,size) */
    int16_t* k;       /* out_size x window */
} coeffs_i16;

static int build_coeffs(uint32_t in_size, uint32_t out_size, coeffs_i16* out) {
    const double support = 3.0;
    double scale = (double)in_size / (double)out_size;
    double filter_scale = scale > 1.0 ? scale : 1.0; /* adaptive kernel size */
    double radius = support * filter_scale;
    uint32_t window = (uint32_t)ceil(radius) * 2 + 1;
    double recip = 1.0 / filter_scale;
    double* wv = (double*)calloc((size_t)window * out_size, sizeof(double));
    out->bounds = (uint32_t*)malloc(sizeof(uint32_t) * 2 * out_size);
    out->k = (int16_t*)calloc((size_t)window * out_size, sizeof(int16_t));
    if (!wv || !out->bounds || !out->k) return -1;
    for (uint32_t o = 0; o < out_size; ++o) {
        double in_center = ((double)o + 0.5) * scale;
        double lo = floor(in_center - radius);
        if (lo < 0.0) lo = 0.0;
        double hi = ceil(in_center + radius);
        if (hi > (double)in_size) hi = (double)in_size;
        uint32_t x_min = (uint32_t)lo, x_max = (uint32_t)hi;
        double center = in_center - 0.5;
        uint32_t bstart = x_min, bend = x_max;
        double* row = wv + (size_t)o * window;
        uint32_t cnt = 0;
        double ww = 0.0;
        for (uint32_t x = x_min; x < x_max; ++x) {
            double w = lanczos3_filter(((double)x - center) * recip);
            if (x == bstart && w == 0.0) {
                bstart++; /* leading zero weights are dropped from the bound */
            } else {
                row[cnt++] = w;
                ww += w;
            }
        }
        for (uint32_t q = cnt; q > 0; --q) { /* trailing zero weights too */
            if (bend <= bstart || row[q - 1] != 0.0) break;
            bend--;
        }
        if (ww != 0.0)
            for (uint32_t q = 0; q < cnt; ++q) row[q] /= ww;
        out->bounds[2 * o] = bstart;
        out->bounds[2 * o + 1] = bend - bstart;
    }
    /* Normalizer16: largest precision such that the max weight still fits an i16 */
    double max_w = 0.0;
    for (size_t q = 0; q < (size_t)window * out_size; ++q)
        if (wv[q] > max_w) max_w = wv[q];
    uint32_t precision = 0;
    for (uint32_t cur = 0; cur < 22; ++cur) { /* PRECISION_BITS = 32 - 8 - 2 */
        precision = cur;
        double nv = round(max_w * (double)(1 << (cur + 1)));
        int32_t next_value = nv >= 2147483647.0 ? 2147483647 : (int32_t)nv;
        if (next_value >= (1 << 15)) break; /* MAX_COEFS_PRECISION = 16 - 1 */
    }
    double sc = (double)(1 << precision);
    for (size_t q = 0; q < (size_t)window * out_size; ++q) {
        double v = round(wv[q] * sc);
        if (v > 32767.0) v = 32767.0;
        if (v < -32768.0) v = -32768.0;
        out->k[q] = (int16_t)v;
    }
    out->window = window;
    out->precision = precision;
    free(wv);
    return 0;
}

int vdfo_resize_coeffs(uint32_t in_size, uint32_t out_size, uint32_t** bounds_out, int16_t** coefs_out,
                       uint32_t* window_out, uint32_t* precision_out) {
    coeffs_i16 c;
    if (build_coeffs(in_size, out_size, &c)) return -1;
    *bounds_out = c.bounds;
    *coefs_out = c.k;
    *window_out = c.window;
    *precision_out = c.precision;
    return 0;
}

static uint8_t clip8(int32_t v, uint32_t precision) {
    int32_t s = v >> precision; /* arithmetic shift, as Rust i32 >> */
    return (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
}

int vdfo_resize_lanczos3(const uint8_t* src, uint32_t w, uint32_t h, size_t pitch, uint32_t left, uint32_t top,
                         uint32_t cw, uint32_t ch, uint32_t out_w, uint32_t out_h, uint8_t* dst) {
    if (cw == 0 || ch == 0 || left + cw > w || top + ch > h) return -1;
    /* the reference materialises the cropped frame (video_hash_builder.rs:198-201) and resizes that with a
     * zero crop box (video_hash.rs:57-59): the window is a standalone image, edges clamp at the crop */
    const uint8_t* base = src + (size_t)top * pitch + left;
    coeffs_i16 hc, vc;
    if (build_coeffs(cw, out_w, &hc)) return -1;
    if (build_coeffs(ch, out_h, &vc)) return -1;
    uint8_t* tmp = (uint8_t*)malloc((size_t)ch * out_w);
    if (!tmp) return -1;
    /* horizontal pass -> u8 temp */
    int32_t init_h = 1 << (hc.precision - 1);
    for (uint32_t y = 0; y < ch; ++y) {
        const uint8_t* row = base + (size_t)y * pitch;
        for (uint32_t o = 0; o < out_w; ++o) {
            const int16_t* k = hc.k + (size_t)o * hc.window;
            uint32_t s0 = hc.bounds[2 * o], sz = hc.bounds[2 * o + 1];
            int32_t ss = init_h;
            for (uint32_t q = 0; q < sz; ++q) ss += (int32_t)row[s0 + q] * (int32_t)k[q];
            tmp[(size_t)y * out_w + o] = clip8(ss, hc.precision);
        }
    }
    /* vertical pass */
    int32_t init_v = 1 << (vc.precision - 1);
    for (uint32_t o = 0; o < out_h; ++o) {
        const int16_t* k = vc.k + (size_t)o * vc.window;
        uint32_t s0 = vc.bounds[2 * o], sz = vc.bounds[2 * o + 1];
        for (uint32_t x = 0; x < out_w; ++x) {
            int32_t ss = init_v;
            for (uint32_t q = 0; q < sz; ++q) ss += (int32_t)tmp[(size_t)(s0 + q) * out_w + x] * (int32_t)k[q];
            dst[(size_t)o * out_w + x] = clip8(ss, vc.precision);
        }
    }
    free(tmp);
    free(hc.bounds), free(hc.k), free(vc.bounds), free(vc.k);
    return 0;
}

/* ---- rustdct 0.7 plan_dct2(16): split-radix DCT-II (one half-size DCT-II on the mirrored sums, two
 * quarter-size DCT-IIs on the rotated mirrored differences), base cases 2 and 4 hard-coded.
 * Twiddles follow rustdct::twiddles::single_twiddle(i, fft_len).conj(). */
static void twiddle(unsigned i, unsigned fft_len, double* re, double* im) {
    double constant = -2.0 * 3.14159265358979323846 / (double)fft_len;
    double angle = constant * (double)i;
    *re = cos(angle);
    *im = -sin(angle);
}

static void dct2_2(double* b) {
    double s = b[0] + b[1];
    b[1] = (b[0] - b[1]) * 0.70710678118654752440; /* FRAC_1_SQRT_2 */
    b[0] = s;
}

static void dct2_4(double* b) {
    double re, im;
    twiddle(1, 16, &re, &im);
    double lower = b[0] - b[3], upper = b[2] - b[1];
    double e[2] = {b[0] + b[3], b[1] + b[2]};
    dct2_2(e);
    b[0] = e[0];
    b[1] = lower * re - upper * im;
    b[2] = e[1];
    b[3] = upper * re + lower * im;
}

static void dct2_split_radix(double* x, unsigned n) {
    if (n == 2) {
        dct2_2(x);
        return;
    }
    if (n == 4) {
        dct2_4(x);
        return;
    }
    unsigned h = n / 2, q = n / 4;
    double d2[8], ev[4], od[4]; /* n <= 16 */
    for (unsigned i = 0; i < q; ++i) {
        double bot = x[i], top = x[n - 1 - i];
        double hb = x[h - 1 - i], ht = x[h + i];
        d2[i] = top + bot;
        d2[h - 1 - i] = hb + ht;
        double lower = bot - top, upper = hb - ht;
        double re, im;
        twiddle(2 * i + 1, 4 * n, &re, &im);
        double c = lower * re + upper * im;
        double s = upper * re - lower * im;
        ev[i] = c;
        od[q - 1 - i] = (i % 2 == 0) ? s : -s;
    }
    dct2_split_radix(d2, h);
    dct2_split_radix(ev, q);
    dct2_split_radix(od, q);
    x[0] = d2[0];
    x[1] = ev[0];
    x[2] = d2[1];
    for (unsigned i = 1; i < q; ++i) {
        double c = ev[i];
        double s = ((i + q) % 2 == 0) ? -od[q - i] : od[q - i];
        x[4 * i - 1] = c + s;
        x[4 * i] = d2[2 * i];
        x[4 * i + 1] = c - s;
        x[4 * i + 2] = d2[2 * i + 1];
    }
    x[n - 1] = -od[0];
}

void vdfo_dct2_16(double* buf) { dct2_split_radix(buf, 16); }

/* raw_dct_ops.rs:107-142 : rows of [t][x][y] (y fastest), then x, then t; transposes are layout only */
void vdfo_dct3d(double* m) {
    double line[16];
    for (int t = 0; t < 16; ++t)
        for (int x = 0; x < 16; ++x) vdfo_dct2_16(m + (t * 16 + x) * 16);
    for (int t = 0; t < 16; ++t)
        for (int y = 0; y < 16; ++y) {
            for (int x = 0; x < 16; ++x) line[x] = m[(t * 16 + x) * 16 + y];
            vdfo_dct2_16(line);
            for (int x = 0; x < 16; ++x) m[(t * 16 + x) * 16 + y] = line[x];
        }
    for (int x = 0; x < 16; ++x)
        for (int y = 0; y < 16; ++y) {
            for (int t = 0; t < 16; ++t) line[t] = m[(t * 16 + x) * 16 + y];
            vdfo_dct2_16(line);
            for (int t = 0; t < 16; ++t) m[(t * 16 + x) * 16 + y] = line[t];
        }
}

void vdfo_hash_from_small(const uint8_t* small, uint64_t hash[16], double* coef_out) {
    double m[4096];
    /* dct_3d.rs:40-44,76 : m[frame][col][row] = pix - 128 */
    for (int t = 0; t < 16; ++t)
        for (int row = 0; row < 16; ++row)
            for (int col = 0; col < 16; ++col) m[(t * 16 + col) * 16 + row] = (double)small[(t * 16 + row) * 16 + col] - 128.0;
    vdfo_dct3d(m);
    if (coef_out) memcpy(coef_out, m, sizeof m);
    memset(hash, 0, 16 * sizeof(uint64_t));
    unsigned bit = 0; /* dct_3d.rs:55-66 row-major over [..10][..10][..10]; video_hash.rs:63-70 Lsb0 */
    for (int t = 0; t < VDFO_HASH_SIZE; ++t)
        for (int x = 0; x < VDFO_HASH_SIZE; ++x)
            for (int y = 0; y < VDFO_HASH_SIZE; ++y, ++bit)
                if (m[(t * 16 + x) * 16 + y] > 0.0) hash[bit / 64] |= (uint64_t)1 << (bit % 64);
}

int vdfo_hash_stack(const uint8_t* frames, uint32_t n_frames, uint32_t w, uint32_t h, size_t pitch,
                    size_t frame_stride, int cropdetect, const uint32_t* frame_dims, uint64_t hash[16],
                    uint32_t crop_out[4], uint8_t* small_out) {
    uint32_t crop[4] = {0, 0, 0, 0};
    memset(hash, 0, 16 * sizeof(uint64_t));
    if (n_frames > VDFO_DCT_SIZE) n_frames = VDFO_DCT_SIZE; /* video_hash_builder.rs:164 take(16) */
    /* are_all_frames_same_size video_hash_builder.rs:169-186 (runs before the frame count is checked) */
    if (frame_dims)
        for (uint32_t f = 0; f + 1 < n_frames; ++f)
            if (frame_dims[2 * f] != frame_dims[2 * f + 2] || frame_dims[2 * f + 1] != frame_dims[2 * f + 3])
                return VDFO_VIDPROC;
    if (n_frames == 0) return VDFO_NOT_ENOUGH_FRAMES; /* detect_crop -> None  :194 */
    if (cropdetect == 1) {
        int st = vdfo_cropdetect_letterbox(frames, n_frames, w, h, pitch, frame_stride, crop);
        if (st) return st;
    }
    if (crop_out) memcpy(crop_out, crop, sizeof crop);
    if (n_frames < VDFO_DCT_SIZE) return VDFO_NOT_ENOUGH_FRAMES; /* dct_3d.rs:47-52, video_hash.rs:61 */
    uint32_t cw = w - crop[0] - crop[1], ch = h - crop[2] - crop[3]; /* crop.rs:92-103 */
    uint8_t small[4096];
    for (uint32_t f = 0; f < 16; ++f)
        if (vdfo_resize_lanczos3(frames + (size_t)f * frame_stride, w, h, pitch, crop[0], crop[2], cw, ch, 16, 16,
                                 small + f * 256))
            return VDFO_VIDPROC;
    if (small_out) memcpy(small_out, small, sizeof small);
    vdfo_hash_from_small(small, hash, NULL);
    return VDFO_OK;
}

int vdfo_hash_stacks(const uint8_t* frames, uint64_t n_stacks, uint32_t w, uint32_t h, int cropdetect,
                     uint64_t* hash_out, int32_t* status_out) {
    size_t fs = (size_t)w * h;
    for (uint64_t s = 0; s < n_stacks; ++s)
        status_out[s] = vdfo_hash_stack(frames + s * 16 * fs, 16, w, h, w, fs, cropdetect, NULL, hash_out + 16 * s,
                                        NULL, NULL);
    return 0;
}
