// vdf.hpp -- C++17 host layer over the C ABI (vdf_b200.h), mirroring the public API of the reference crate
// (vid_dup_finder_lib/src/lib.rs:132-140): VideoHash, VideoHashBuilder / CreationOptions, search,
// search_with_references, MatchGroup, Error, Cropdetect, DEFAULT_*.  The reference is compiled Rust; no Rust
// toolchain exists in the build image, so this header is the compiled-language host side (INTEGRATION.md shows the
// equivalent Rust `extern "C"` binding).  Same names, argument meaning and error behaviour as the crate; what runs
// here is exactly what the boundary leaves to the host: the stable (duration, path) sort, the tolerance cast,
// index -> path, MatchGroup construction/validation.  All comparison, grouping and hashing work is done by the CUDA
// library; there is no CPU fallback (a missing GPU surfaces as vdf::DeviceError).
#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "vdf_b200.h"

namespace vdf {

// ---- definitions.rs ---------------------------------------------------------------------------------------
constexpr double DEFAULT_SEARCH_TOLERANCE = 0.35;       // definitions.rs:5
constexpr double DEFAULT_VID_HASH_SKIP_FORWARD = 15.0;  // definitions.rs:18
constexpr double DEFAULT_VID_HASH_DURATION = 10.0;      // definitions.rs:29
constexpr uint32_t DCT_SIZE = 16, HASH_SIZE = 10, HASH_BITS = 1000, HASH_WORDS = 16;  // definitions.rs:34-43
constexpr double TOLERANCE_SCALING_FACTOR = 1000.0;     // definitions.rs:40

enum class Cropdetect { None, Letterbox, Motion };  // definitions.rs:46-54

// (tolerance * TOLERANCE_SCALING_FACTOR) as u32 -- search_algorithm.rs:82 (truncating, saturating, NaN -> 0)
inline uint32_t tolerance_to_int(double tolerance) {
    const double v = tolerance * TOLERANCE_SCALING_FACTOR;
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

// ---- errors -----------------------------------------------------------------------------------------------
struct DeviceError : std::runtime_error {  // not part of the crate: the GPU library failed (no CPU fallback)
    int code;
    DeviceError(int c, const std::string& m) : std::runtime_error("vdf_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

struct Error {  // vid_dup_finder_lib::Error, video_hashing/mod.rs:17-28
    enum Kind { NotVideo, VidProc, NotEnoughFrames } kind;
    std::string detail;
    std::string to_string() const {
        switch (kind) {
            case NotVideo: return "File is not a video";
            case VidProc: return "Video processing error: " + detail;
            default: return "Could not extract enough frames";
        }
    }
};

// ---- context ----------------------------------------------------------------------------------------------
class Context {  // one GPU and one stream -- or several GPUs of one node behind one context -- not thread-safe (vdf_b200.h)
  public:
    explicit Context(int device = 0) {
        int rc = vdf_ctx_create(device, &h_);
        if (rc != VDF_OK) throw DeviceError(rc, "vdf_ctx_create failed (no sm_100 device?)");
    }
    // search(), search_with_references() and VideoHashBuilder then use every listed GPU (vdf_ctx_create_multi)
    explicit Context(const std::vector<int>& devices) {
        int rc = vdf_ctx_create_multi(devices.data(), (int)devices.size(), &h_);
        if (rc != VDF_OK) throw DeviceError(rc, "vdf_ctx_create_multi failed (sm_100 devices with peer access?)");
    }
    int device_count() const { return vdf_ctx_device_count(h_); }
    ~Context() { vdf_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    vdf_ctx* get() const { return h_; }
    void check(int rc) const {
        if (rc != VDF_OK) throw DeviceError(rc, vdf_last_error(h_));
    }

  private:
    vdf_ctx* h_ = nullptr;
};

// ---- Rust Path ordering (Unix): component-wise, RootDir < CurDir < ParentDir < Normal(bytes) -----------------
namespace detail {
struct Comp {
    int kind;
    const char* s;
    size_t n;
};
inline std::vector<Comp> components(const std::string& p) {
    std::vector<Comp> out;
    size_t i = 0;
    const bool root = !p.empty() && p[0] == '/';
    if (root) out.push_back({1, p.data(), 1}), i = 1;
    bool first = true;
    while (i <= p.size()) {
        size_t j = p.find('/', i);
        if (j == std::string::npos) j = p.size();
        const size_t n = j - i;
        if (n == 1 && p[i] == '.') {
            if (first && !root && i == 0) out.push_back({2, p.data(), 1});
        } else if (n > 0) {
            out.push_back({(n == 2 && p[i] == '.' && p[i + 1] == '.') ? 3 : 4, p.data() + i, n});
        }
        first = false;
        i = j + 1;
    }
    return out;
}
}  // namespace detail

inline int path_cmp(const std::string& a, const std::string& b) {
    const auto ca = detail::components(a), cb = detail::components(b);
    for (size_t k = 0; k < ca.size() && k < cb.size(); ++k) {
        if (ca[k].kind != cb[k].kind) return ca[k].kind < cb[k].kind ? -1 : 1;
        if (ca[k].kind == 4) {
            const int c = std::memcmp(ca[k].s, cb[k].s, std::min(ca[k].n, cb[k].n));
            if (c) return c < 0 ? -1 : 1;
            if (ca[k].n != cb[k].n) return ca[k].n < cb[k].n ? -1 : 1;
        }
    }
    return ca.size() == cb.size() ? 0 : (ca.size() < cb.size() ? -1 : 1);
}

// ---- VideoHash (video_hash.rs:26-32) ------------------------------------------------------------------------
class VideoHash {
  public:
    VideoHash() = default;
    VideoHash(const std::array<uint64_t, HASH_WORDS>& words, std::string src_path, uint32_t duration)
        : hash_(words), src_path_(std::move(src_path)), duration_(duration) {}
    const std::string& src_path() const { return src_path_; }  // video_hash.rs:178-181
    uint32_t duration() const { return duration_; }            // video_hash.rs:184-187
    // video_hash.rs:190-192,311-317: a scalar accessor (callers sort single pairs with it); bulk work is the GPU's
    uint32_t hamming_distance(const VideoHash& o) const {
        uint32_t acc = 0;
        for (uint32_t w = 0; w < HASH_WORDS; ++w) acc += (uint32_t)__builtin_popcountll(hash_[w] ^ o.hash_[w]);
        return acc;
    }
    double normalized_hamming_distance(const VideoHash& o) const { return hamming_distance(o) / TOLERANCE_SCALING_FACTOR; }
    const std::array<uint64_t, HASH_WORDS>& words() const { return hash_; }
    std::vector<bool> raw_hash() const {  // video_hash.rs:207-216
        std::vector<bool> b(HASH_BITS);
        for (uint32_t k = 0; k < HASH_BITS; ++k) b[k] = (hash_[k / 64] >> (k % 64)) & 1;
        return b;
    }
    // test_util (video_hash.rs:240-308)
    VideoHash with_duration(uint32_t d) const { return VideoHash(hash_, src_path_, d); }
    VideoHash with_src_path(std::string p) const { return VideoHash(hash_, std::move(p), duration_); }
    static VideoHash full_hash(std::string name) {
        std::array<uint64_t, HASH_WORDS> w;
        w.fill(~0ull);
        return VideoHash(w, std::move(name), 0);
    }
    static VideoHash empty_hash(std::string name) { return VideoHash({}, std::move(name), 0); }
    template <class Rng>
    static VideoHash random_hash(Rng& rng) {
        std::array<uint64_t, HASH_WORDS> w;
        for (auto& x : w) x = rng();
        w[15] &= (1ull << 40) - 1;
        return VideoHash(w, "", 0);
    }
    // a hash exactly `target` bits away (any of the 1024 stored bits), video_hash.rs:272-291
    template <class Rng>
    VideoHash hash_with_spatial_distance(uint32_t target, Rng& rng) const {
        std::array<uint64_t, HASH_WORDS> w = hash_;
        std::array<uint16_t, 1024> pos;
        for (int k = 0; k < 1024; ++k) pos[k] = (uint16_t)k;
        for (uint32_t k = 0; k < target; ++k) {
            std::swap(pos[k], pos[k + rng() % (1024 - k)]);
            w[pos[k] / 64] ^= 1ull << (pos[k] % 64);
        }
        return VideoHash(w, src_path_, duration_);
    }
    bool operator==(const VideoHash& o) const { return hash_ == o.hash_ && src_path_ == o.src_path_ && duration_ == o.duration_; }

  private:
    std::array<uint64_t, HASH_WORDS> hash_{};
    std::string src_path_;
    uint32_t duration_ = 0;
};

// ---- MatchGroup (matches/match_group.rs) ----------------------------------------------------------------------
struct TooFewEntries {};

class MatchGroup {
  public:
    static std::variant<MatchGroup, TooFewEntries> create(std::vector<std::string> entries) {  // MatchGroup::new :21-30
        if (entries.size() < 2) return TooFewEntries{};
        return MatchGroup(std::nullopt, std::move(entries));
    }
    static std::variant<MatchGroup, TooFewEntries> create_with_reference(std::string reference,
                                                                         std::vector<std::string> entries) {  // :35-47
        if (entries.empty()) return TooFewEntries{};
        return MatchGroup(std::move(reference), std::move(entries));
    }
    size_t len() const { return duplicates_.size(); }                                  // :51-53
    const std::optional<std::string>& reference() const { return reference_; }           // :57-59
    const std::vector<std::string>& duplicates() const { return duplicates_; }           // :62-64
    std::vector<std::string> contained_paths() const {                                   // :68-82
        std::vector<std::string> v = duplicates_;
        if (reference_) v.push_back(*reference_);
        return v;
    }
    std::vector<MatchGroup> dup_combinations() const {  // :89-105
        std::vector<MatchGroup> out;
        if (reference_) {
            for (const auto& d : duplicates_) out.push_back(MatchGroup(*reference_, {d}));
        } else {
            for (size_t a = 0; a < duplicates_.size(); ++a)
                for (size_t b = a + 1; b < duplicates_.size(); ++b) out.push_back(MatchGroup(std::nullopt, {duplicates_[a], duplicates_[b]}));
        }
        return out;
    }
    bool operator==(const MatchGroup& o) const { return reference_ == o.reference_ && duplicates_ == o.duplicates_; }

  private:
    MatchGroup(std::optional<std::string> r, std::vector<std::string> d) : reference_(std::move(r)), duplicates_(std::move(d)) {}
    std::optional<std::string> reference_;
    std::vector<std::string> duplicates_;
};

// ---- search (video_dup_finder.rs) -----------------------------------------------------------------------------
namespace detail {
inline int cropdetect_code(Cropdetect c) {  // definitions.rs:46-54 -> VDF_CROPDETECT_*
    return c == Cropdetect::None ? VDF_CROPDETECT_NONE : c == Cropdetect::Letterbox ? VDF_CROPDETECT_LETTERBOX : VDF_CROPDETECT_MOTION;
}
// Vec<VideoHash> -> the struct-of-arrays the C ABI takes (caller's order; paths as one blob + n+1 offsets)
struct Soa {
    std::vector<uint64_t> words;
    std::vector<uint32_t> dur;
    std::vector<char> blob;
    std::vector<uint64_t> off;
    explicit Soa(const std::vector<VideoHash>& v, bool with_paths = true) {
        words.resize(v.size() * HASH_WORDS);
        dur.resize(v.size());
        off.assign(v.size() + 1, 0);
        size_t total = 0;
        for (size_t k = 0; k < v.size(); ++k) {
            std::memcpy(&words[k * HASH_WORDS], v[k].words().data(), HASH_WORDS * 8);
            dur[k] = v[k].duration();
            if (with_paths) total += v[k].src_path().size();
            off[k + 1] = total;
        }
        blob.resize(total + 1);
        if (with_paths)
            for (size_t k = 0; k < v.size(); ++k) std::memcpy(blob.data() + off[k], v[k].src_path().data(), v[k].src_path().size());
    }
};
// Search::sort, search_algorithm.rs:55-61 (stable, key (duration, src_path)) as the library computes it
inline std::vector<uint64_t> sort_order(const std::vector<VideoHash>& v) {
    Soa soa(v);
    std::vector<uint64_t> order(v.size());
    if (vdf_sort_order(soa.dur.data(), soa.blob.data(), soa.off.data(), v.size(), order.data()) != VDF_OK)
        throw DeviceError(VDF_ERR_INVALID, "vdf_sort_order");
    return order;
}
}  // namespace detail

// video_dup_finder.rs:7-13
inline std::vector<MatchGroup> search(const std::vector<VideoHash>& hashes, double tolerance, Context& ctx) {
    std::vector<MatchGroup> out;
    if (hashes.empty()) return out;  // search_algorithm.rs:88-90
    detail::Soa soa(hashes);         // the caller's order: the library sorts (vdf_search, csrc/host.cu)
    vdf_groups g{};
    ctx.check(vdf_search(ctx.get(), soa.words.data(), soa.dur.data(), soa.blob.data(), soa.off.data(), hashes.size(), tolerance, &g));
    out.reserve(g.n_groups);
    for (uint64_t k = 0; k < g.n_groups; ++k) {
        std::vector<std::string> paths;
        for (uint64_t m = g.group_ptr[k]; m < g.group_ptr[k + 1]; ++m) paths.push_back(hashes[g.member_idx[m]].src_path());
        auto mg = MatchGroup::create(std::move(paths));  // .filter_map(|x| MatchGroup::new(x).ok())
        if (auto* ok = std::get_if<MatchGroup>(&mg)) out.push_back(std::move(*ok));
    }
    vdf_free_groups(&g);
    return out;
}

// video_dup_finder.rs:19-46
inline std::vector<MatchGroup> search_with_references(const std::vector<VideoHash>& ref_hashes,
                                                      const std::vector<VideoHash>& new_hashes, double tolerance, Context& ctx) {
    std::vector<MatchGroup> out;
    if (ref_hashes.empty() || new_hashes.empty()) return out;
    detail::Soa refs(ref_hashes, false), cands(new_hashes);  // references keep the caller's order and need no paths
    vdf_csr c{};
    ctx.check(vdf_search_with_references(ctx.get(), refs.words.data(), refs.dur.data(), ref_hashes.size(), cands.words.data(),
                                         cands.dur.data(), cands.blob.data(), cands.off.data(), new_hashes.size(), tolerance, &c));
    for (uint64_t r = 0; r < c.n_rows; ++r) {
        if (c.row_ptr[r + 1] == c.row_ptr[r]) continue;  // video_dup_finder.rs:38-43
        std::vector<std::string> paths;
        for (uint64_t m = c.row_ptr[r]; m < c.row_ptr[r + 1]; ++m) paths.push_back(new_hashes[c.col_idx[m]].src_path());
        auto mg = MatchGroup::create_with_reference(ref_hashes[r].src_path(), std::move(paths));
        if (auto* ok = std::get_if<MatchGroup>(&mg)) out.push_back(std::move(*ok));
    }
    vdf_free_csr(&c);
    return out;
}

// `Search` (search_algorithm.rs:5-53,81-198): `Search::from(hashes)` sorts once and keeps the table -- here resident in HBM
// in the kernels' layout (vdf_table) -- and `search_self(tolerance)` runs on it any number of times (a tolerance sweep, the
// app's repeated searches over one loaded cache).  Returned groups hold paths, in the reference's order.
class Search {
  public:
    Search(const std::vector<VideoHash>& hashes, Context& ctx) : ctx_(ctx) {
        const std::vector<uint64_t> order = detail::sort_order(hashes);  // Search::sort (:55-61)
        std::vector<uint64_t> words(hashes.size() * HASH_WORDS);
        std::vector<uint32_t> dur(hashes.size());
        paths_.reserve(hashes.size());
        for (size_t k = 0; k < order.size(); ++k) {
            const VideoHash& h = hashes[order[k]];
            std::memcpy(&words[k * HASH_WORDS], h.words().data(), HASH_WORDS * 8);
            dur[k] = h.duration();
            paths_.push_back(h.src_path());
        }
        ctx_.check(vdf_table_create(ctx_.get(), words.data(), dur.data(), hashes.size(), &t_));
    }
    ~Search() { vdf_table_destroy(t_); }
    Search(const Search&) = delete;
    Search& operator=(const Search&) = delete;
    size_t len() const { return paths_.size(); }

    std::vector<std::vector<std::string>> search_self(double tolerance) {  // :81-171
        std::vector<std::vector<std::string>> out;
        vdf_groups g{};
        ctx_.check(vdf_table_search_self_groups(t_, tolerance_to_int(tolerance), &g));
        out.reserve(g.n_groups);
        for (uint64_t k = 0; k < g.n_groups; ++k) {
            std::vector<std::string> paths;
            for (uint64_t m = g.group_ptr[k]; m < g.group_ptr[k + 1]; ++m) paths.push_back(paths_[g.member_idx[m]]);
            out.push_back(std::move(paths));
        }
        vdf_free_groups(&g);
        return out;
    }

  private:
    Context& ctx_;
    vdf_table* t_ = nullptr;
    std::vector<std::string> paths_;
};

// ---- the app's hash cache file (format: see vdf_cache_load in vdf_b200.h) ------------------------
// the Ok(VideoHash) entries of a cache file, in file order -- what the app feeds `search` after loading its cache
// (app_fns.rs:428-482); cached errors are skipped
inline std::vector<VideoHash> load_hash_cache(const std::string& file) {
    vdf_cache c{};
    const int rc = vdf_cache_load(file.c_str(), &c);
    if (rc != VDF_OK) throw DeviceError(rc, "cannot read hash cache " + file);
    std::vector<VideoHash> out;
    for (uint64_t i = 0; i < c.n; ++i) {
        if (c.kind[i] != VDF_CACHE_OK) continue;
        std::array<uint64_t, HASH_WORDS> w;
        std::memcpy(w.data(), c.hashes + i * HASH_WORDS, HASH_WORDS * 8);
        out.emplace_back(w, std::string(c.src_blob + c.src_off[i], c.src_blob + c.src_off[i + 1]), c.durations[i]);
    }
    vdf_free_cache(&c);
    return out;
}

// ---- hashing (video_hash_builder.rs) ----------------------------------------------------------------------------
struct CreationOptions {  // video_hash_builder.rs:17-63
    double skip_forward_amount = DEFAULT_VID_HASH_SKIP_FORWARD;
    double duration = DEFAULT_VID_HASH_DURATION;
    Cropdetect cropdetect = Cropdetect::Letterbox;
};

struct GrayFrame {  // one decoded frame (image::GrayImage): row-major u8
    const uint8_t* data;
    uint32_t width, height, pitch;
};

using HashResult = std::variant<VideoHash, Error>;

class VideoHashBuilder {  // video_hash_builder.rs:70-83; decoding (:85-167) stays with the caller
  public:
    explicit VideoHashBuilder(Context& ctx, CreationOptions options = {}) : ctx_(ctx), options_(options) {}
    static VideoHashBuilder from_options(Context& ctx, CreationOptions options) { return VideoHashBuilder(ctx, options); }

    // gen_hash's compute tail for one video (video_hash_builder.rs:214-223 after iterate_video_frames)
    HashResult hash_frames(const std::vector<GrayFrame>& frames, const std::string& src_path, uint32_t duration_secs) {
        return hash_many({frames}, {src_path}, {duration_secs})[0];
    }

    // many videos per launch: the frames of each stack are gathered into one staging buffer
    std::vector<HashResult> hash_many(const std::vector<std::vector<GrayFrame>>& stacks, const std::vector<std::string>& paths,
                                      const std::vector<uint32_t>& durations) {
        const uint32_t n = (uint32_t)stacks.size();
        std::vector<vdf_stack_desc> desc(n);
        std::vector<uint8_t> buf;
        for (uint32_t s = 0; s < n; ++s) {
            vdf_stack_desc& d = desc[s];
            std::memset(&d, 0, sizeof d);
            const size_t nf = std::min<size_t>(stacks[s].size(), DCT_SIZE);  // take(DCT_SIZE), :164
            d.n_frames = (uint32_t)nf;
            if (nf == 0) continue;
            const GrayFrame& f0 = stacks[s][0];
            bool same = true;
            for (size_t f = 1; f < nf; ++f) same &= stacks[s][f].width == f0.width && stacks[s][f].height == f0.height;
            if (!same) {  // are_all_frames_same_size, :169-186
                d.flags = VDF_STACK_FLAG_MIXED_SIZES;
                continue;
            }
            d.offset = buf.size(), d.width = f0.width, d.height = f0.height, d.pitch = f0.width;
            d.frame_stride = (uint64_t)f0.width * f0.height;
            for (size_t f = 0; f < nf; ++f)
                for (uint32_t y = 0; y < f0.height; ++y)
                    buf.insert(buf.end(), stacks[s][f].data + (size_t)y * stacks[s][f].pitch,
                               stacks[s][f].data + (size_t)y * stacks[s][f].pitch + f0.width);
        }
        if (buf.empty()) buf.push_back(0);
        std::vector<uint64_t> words((size_t)n * HASH_WORDS);
        std::vector<int32_t> status(n);
        ctx_.check(vdf_hash_stacks(ctx_.get(), buf.data(), desc.data(), n,
                                   detail::cropdetect_code(options_.cropdetect),
                                   words.data(), status.data(), nullptr));
        std::vector<HashResult> out;
        for (uint32_t s = 0; s < n; ++s) {
            if (status[s] == VDF_STACK_OK) {
                std::array<uint64_t, HASH_WORDS> w;
                std::memcpy(w.data(), &words[(size_t)s * HASH_WORDS], sizeof w);
                out.emplace_back(VideoHash(w, paths[s], durations[s]));
            } else if (status[s] == VDF_STACK_VIDPROC) {
                out.emplace_back(Error{Error::VidProc, "frames not all same size"});
            } else {
                out.emplace_back(Error{Error::NotEnoughFrames, ""});
            }
        }
        return out;
    }

  private:
    Context& ctx_;
    CreationOptions options_;
};

// Batch-oriented hashing: the per-file rayon loop of the app (video_hash_filesystem_cache.rs:237-257)
// becomes "decode threads push(), a collector takes results"; the library batches (vdf_pipeline_*, csrc/pipeline.cu).
// push() is thread-safe and copies the frames into pinned memory before returning.
class HashPipeline {
  public:
    struct Item {
        std::string src_path;
        HashResult result;
    };
    HashPipeline(Context& ctx, CreationOptions options = {}, uint32_t max_batch_stacks = 64, uint64_t batch_bytes = 1ull << 30) {
        const int rc = vdf_pipeline_create(ctx.get(), max_batch_stacks, batch_bytes, detail::cropdetect_code(options.cropdetect), &p_);
        if (rc != VDF_OK) throw DeviceError(rc, "vdf_pipeline_create");
    }
    ~HashPipeline() { vdf_pipeline_destroy(p_); }
    HashPipeline(const HashPipeline&) = delete;
    HashPipeline& operator=(const HashPipeline&) = delete;

    void push(const std::vector<GrayFrame>& frames, std::string src_path, uint32_t duration_secs) {
        const size_t nf = std::min<size_t>(frames.size(), DCT_SIZE);  // take(DCT_SIZE), video_hash_builder.rs:164
        std::vector<const uint8_t*> ptr(nf ? nf : 1, nullptr);
        uint32_t flags = 0, w = 0, h = 0, pitch = 0;
        for (size_t f = 0; f < nf; ++f) {
            ptr[f] = frames[f].data;
            if (f == 0) w = frames[f].width, h = frames[f].height, pitch = frames[f].pitch;
            else if (frames[f].width != w || frames[f].height != h || frames[f].pitch != pitch) {
                if (frames[f].width != w || frames[f].height != h) flags = VDF_STACK_FLAG_MIXED_SIZES;  // :169-186
                else throw std::invalid_argument("HashPipeline::push: frames of one stack must share a pitch");
            }
        }
        uint64_t tag;
        {
            std::lock_guard<std::mutex> lk(m_);
            tag = meta_.size();
            meta_.emplace_back(std::move(src_path), duration_secs);
        }
        const int rc = vdf_pipeline_push(p_, tag, ptr.data(), (uint32_t)nf, w, h, pitch, flags);
        if (rc != VDF_OK) throw DeviceError(rc, vdf_pipeline_error(p_));
    }
    void flush() {
        const int rc = vdf_pipeline_flush(p_);
        if (rc != VDF_OK) throw DeviceError(rc, vdf_pipeline_error(p_));
    }
    // finished videos, in completion order; wait = block until at least one is ready (or nothing is in flight)
    std::vector<Item> results(bool wait = false) {
        std::vector<vdf_pipeline_result> buf(1024);
        uint32_t n = 0;
        const int rc = vdf_pipeline_poll(p_, buf.data(), (uint32_t)buf.size(), &n, wait ? 1 : 0);
        if (rc != VDF_OK) throw DeviceError(rc, vdf_pipeline_error(p_));
        std::vector<Item> out;
        std::lock_guard<std::mutex> lk(m_);
        for (uint32_t k = 0; k < n; ++k) {
            const auto& meta = meta_[buf[k].tag];
            if (buf[k].status == VDF_STACK_OK) {
                std::array<uint64_t, HASH_WORDS> w;
                std::memcpy(w.data(), buf[k].hash, sizeof w);
                out.push_back({meta.first, VideoHash(w, meta.first, meta.second)});
            } else if (buf[k].status == VDF_STACK_VIDPROC) {
                out.push_back({meta.first, Error{Error::VidProc, "frames not all same size"}});
            } else {
                out.push_back({meta.first, Error{Error::NotEnoughFrames, ""}});
            }
        }
        return out;
    }

  private:
    vdf_hash_pipeline* p_ = nullptr;
    std::mutex m_;
    std::vector<std::pair<std::string, uint32_t>> meta_;
};

}  // namespace vdf
