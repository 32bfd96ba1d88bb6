// C++ re-expression of the reference's search tests against the compiled-language host layer (include/vdf.hpp):
//   vid_dup_finder_lib/tests/test_find_all.rs:137-315 (4 tests), search_algorithm.rs:203-208, video_hash.rs:325-371.
// Build + run: see tests/test_cpp_host_layer.py (needs a B200: the host layer has no CPU fallback).
#include <cstdio>
#include <random>
#include <thread>

#include "vdf.hpp"

using namespace vdf;

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            return 1;                                                        \
        }                                                                    \
    } while (0)

using Rng = std::mt19937_64;

// test_find_all.rs:14-66
struct HashesWithDistance {
    VideoHash start_hash;
    std::vector<VideoHash> members;
    HashesWithDistance(VideoHash start, uint32_t distance, uint32_t num, Rng& rng) : start_hash(std::move(start)) {
        for (uint32_t i = 0; i < num; ++i) members.push_back(start_hash.hash_with_spatial_distance(distance, rng));
    }
    std::vector<VideoHash> shuffled(Rng& rng) const {
        auto v = members;
        std::shuffle(v.begin(), v.end(), rng);
        return v;
    }
};

// test_find_all.rs:68-132
struct HashesWithDistanceSet {
    std::vector<HashesWithDistance> groups;
    HashesWithDistanceSet(uint32_t num_groups, uint32_t per_group, uint32_t inter, uint32_t intra, Rng& rng) {
        VideoHash start = VideoHash::random_hash(rng);
        uint32_t cur = 0;
        for (uint32_t g = 0; g < num_groups; ++g) {
            groups.emplace_back(start.hash_with_spatial_distance(cur, rng), intra, per_group, rng);
            cur += inter;
            per_group += 10;
        }
    }
    std::vector<VideoHash> all_members(Rng& rng) const {
        std::vector<VideoHash> v;
        for (const auto& g : groups)
            for (auto& m : g.shuffled(rng)) v.push_back(m);
        std::shuffle(v.begin(), v.end(), rng);
        return v;
    }
};

static void name_all(std::vector<VideoHash>& v, const char* prefix = "p") {
    for (size_t i = 0; i < v.size(); ++i) v[i] = v[i].with_src_path(std::string(prefix) + "/" + std::to_string(1000000 + i));
}

static int test_metric(Context&) {  // video_hash.rs:325-371
    Rng rng(1);
    for (int i = 0; i < 1000; ++i) {
        auto a = VideoHash::random_hash(rng), b = VideoHash::random_hash(rng), c = VideoHash::random_hash(rng);
        CHECK(a.hamming_distance(b) <= a.hamming_distance(c) + b.hamming_distance(c));
        CHECK(a.hamming_distance(b) == b.hamming_distance(a));
    }
    CHECK(VideoHash::empty_hash("").hamming_distance(VideoHash::empty_hash("")) == 0);
    CHECK(VideoHash::full_hash("").hamming_distance(VideoHash::full_hash("")) == 0);
    CHECK(VideoHash::full_hash("").hamming_distance(VideoHash::empty_hash("")) == 1024);
    return 0;
}

static int test_searching_nothing_returns_empty_vec(Context& ctx) {  // search_algorithm.rs:203-208
    CHECK(search({}, 1.0, ctx).empty());
    return 0;
}

static int test_find_dups_finds_a_known_group(Context& ctx) {  // test_find_all.rs:137-169
    Rng rng(1);
    HashesWithDistanceSet groups(1, 50, 201, 100, rng);
    auto members = groups.all_members(rng);
    name_all(members);
    auto dups = search(members, 200 / TOLERANCE_SCALING_FACTOR, ctx);
    CHECK(dups.size() == 1);
    CHECK(dups[0].len() == 50);
    return 0;
}

static int test_find_dups_discriminates_by_duration(Context& ctx) {  // test_find_all.rs:176-238
    Rng rng(2);
    HashesWithDistanceSet groups(1, 100, 201, 100, rng);
    auto short_group = groups.groups[0].shuffled(rng);
    for (size_t i = 0; i < short_group.size(); ++i) short_group[i] = short_group[i].with_duration(50).with_src_path("short/" + std::to_string(i));
    std::vector<VideoHash> all = short_group;
    for (size_t i = 0; i < 50; ++i) all.push_back(short_group[i].with_duration(250).with_src_path("long/" + std::to_string(i)));
    std::shuffle(all.begin(), all.end(), rng);
    auto dups = search(all, 200 / TOLERANCE_SCALING_FACTOR, ctx);
    std::sort(dups.begin(), dups.end(), [](const MatchGroup& a, const MatchGroup& b) { return a.len() < b.len(); });
    CHECK(dups.size() == 2);
    CHECK(dups[1].len() == 100);
    CHECK(dups[0].len() == 50);
    for (const auto& p : dups[0].duplicates()) CHECK(p.rfind("long/", 0) == 0);
    for (const auto& p : dups[1].duplicates()) CHECK(p.rfind("short/", 0) == 0);
    return 0;
}

static int test_find_dups_discriminates_by_distance(Context& ctx) {  // test_find_all.rs:244-269
    Rng rng(3);
    HashesWithDistanceSet groups(2, 100, 150, 50, rng);
    auto all = groups.all_members(rng);
    name_all(all);
    auto dups = search(all, 100 / TOLERANCE_SCALING_FACTOR, ctx);
    std::sort(dups.begin(), dups.end(), [](const MatchGroup& a, const MatchGroup& b) { return a.len() < b.len(); });
    CHECK(dups.size() == 2);
    CHECK(dups[0].len() == 100);
    CHECK(dups[1].len() == 110);
    return 0;
}

static int test_find_with_refs(Context& ctx) {  // test_find_all.rs:273-315
    Rng rng(4);
    HashesWithDistanceSet groups(5, 100, 150, 50, rng);
    auto cands = groups.all_members(rng);
    name_all(cands);
    CHECK(cands.size() == 100 + 110 + 120 + 130 + 140);
    auto dups = search_with_references({groups.groups[3].start_hash.with_src_path("ref3")}, cands, 50 / TOLERANCE_SCALING_FACTOR, ctx);
    CHECK(dups.size() == 1);
    CHECK(dups[0].len() == 130);
    CHECK(dups[0].reference().value() == "ref3");
    auto dups2 = search_with_references({groups.groups[0].start_hash.with_src_path("ref0"), groups.groups[4].start_hash.with_src_path("ref4")},
                                        cands, 50 / TOLERANCE_SCALING_FACTOR, ctx);
    CHECK(dups2.size() == 2);
    CHECK(dups2[0].len() == 100);
    CHECK(dups2[1].len() == 140);
    return 0;
}

static int test_group_order_and_paths(Context& ctx) {
    // three identical hashes + a chain: groups list matches in sorted (duration, path) order, the target last,
    // groups in descending target order (search_algorithm.rs:136,158-161,167); path order is component-wise
    auto e = VideoHash::empty_hash("");
    std::vector<VideoHash> v = {e.with_src_path("a/b"), e.with_src_path("a-b"), e.with_src_path("a")};
    auto g = search(v, 0.0, ctx);
    CHECK(g.size() == 1);
    CHECK((g[0].duplicates() == std::vector<std::string>{"a/b", "a-b", "a"}));  // sorted: a, a/b, a-b ; target "a" last
    CHECK(path_cmp("a//b/", "a/b") == 0 && path_cmp("./a", "a") < 0 && path_cmp("/a", "a") < 0 && path_cmp("a", "a/b") < 0);
    CHECK(tolerance_to_int(0.35) == 350 && tolerance_to_int(-1) == 0 && tolerance_to_int(1e12) == 4294967295u);
    return 0;
}

static int test_hash_frames(Context& ctx) {
    // 16 frames of a moving gradient; a second video is the same content letterboxed: both must hash, and match
    const uint32_t W = 160, H = 90;
    std::vector<std::vector<uint8_t>> a(16, std::vector<uint8_t>(W * H)), b(16, std::vector<uint8_t>((W + 40) * (H + 30), 16));
    for (uint32_t t = 0; t < 16; ++t)
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) {
                uint8_t p = (uint8_t)(40 + ((x * 3 + y * 2 + t * 11) % 160));
                a[t][y * W + x] = p;
                b[t][(y + 15) * (W + 40) + x + 20] = p;
            }
    std::vector<GrayFrame> fa, fb;
    for (uint32_t t = 0; t < 16; ++t) fa.push_back({a[t].data(), W, H, W}), fb.push_back({b[t].data(), W + 40, H + 30, W + 40});
    VideoHashBuilder builder(ctx);
    auto ra = builder.hash_frames(fa, "a.mp4", 30), rb = builder.hash_frames(fb, "b.mp4", 31);
    CHECK(std::holds_alternative<VideoHash>(ra) && std::holds_alternative<VideoHash>(rb));
    CHECK(std::get<VideoHash>(ra).hamming_distance(std::get<VideoHash>(rb)) == 0);  // the crop removes the bars exactly
    auto groups = search({std::get<VideoHash>(ra), std::get<VideoHash>(rb)}, DEFAULT_SEARCH_TOLERANCE, ctx);
    CHECK(groups.size() == 1 && groups[0].len() == 2);
    {  // the same two videos (plus a short one) through the batch pipeline, pushed from two threads
        HashPipeline pipe(ctx, {}, 2, (uint64_t)16 * (W + 40) * (H + 30) * 2 + 4096);
        std::thread t1([&] { pipe.push(fa, "a.mp4", 30); pipe.push(std::vector<GrayFrame>(fa.begin(), fa.begin() + 5), "short.mp4", 1); });
        std::thread t2([&] { pipe.push(fb, "b.mp4", 31); });
        t1.join(), t2.join();
        pipe.flush();
        auto items = pipe.results();
        CHECK(items.size() == 3);
        for (const auto& it : items) {
            if (it.src_path == "short.mp4") CHECK(std::holds_alternative<Error>(it.result) && std::get<Error>(it.result).kind == Error::NotEnoughFrames);
            else CHECK(std::holds_alternative<VideoHash>(it.result) && std::get<VideoHash>(it.result).hamming_distance(std::get<VideoHash>(ra)) == 0);
        }
    }
    fa.pop_back();
    auto rc = builder.hash_frames(fa, "c.mp4", 30);
    CHECK(std::holds_alternative<Error>(rc) && std::get<Error>(rc).kind == Error::NotEnoughFrames);
    fb[3] = fa[3];
    auto rd = builder.hash_frames(fb, "d.mp4", 30);
    CHECK(std::holds_alternative<Error>(rd) && std::get<Error>(rd).kind == Error::VidProc);
    return 0;
}

// host-only: Search::sort as the library computes it (vdf_sort_order, multi-threaded) against a plain stable sort with the
// component-wise comparator of this header; needs no GPU
// `Search::from` once, `search_self` at several tolerances: the prepared table gives what the one-shot `search` gives
static int test_prepared_search_object(Context& ctx) {
    Rng rng(5);
    HashesWithDistanceSet sets(3, 60, 200, 40, rng);
    auto all = sets.all_members(rng);
    name_all(all);
    Search s(all, ctx);
    CHECK(s.len() == all.size());
    for (double tol : {0.0, 0.05, 0.1, 0.35}) {
        const auto a = s.search_self(tol);
        const auto b = search(all, tol, ctx);
        CHECK(a.size() == b.size());
        for (size_t k = 0; k < a.size(); ++k) CHECK(a[k] == b[k].duplicates());
    }
    return 0;
}

// one context over every GPU of the box (vdf_ctx_create_multi): the same calls, the same results; skipped below 2 GPUs
static int test_multi_device_context(Context& ctx) {
    int n_dev = 0;
    for (int d = 0; d < 8; ++d) {  // probe by creating contexts: the test links against the C ABI only
        vdf_ctx* p = nullptr;
        if (vdf_ctx_create(d, &p) != VDF_OK) break;
        vdf_ctx_destroy(p);
        ++n_dev;
    }
    if (n_dev < 2) {
        std::printf("       (skipped: %d GPU)\n", n_dev);
        return 0;
    }
    std::vector<int> devs(n_dev);
    for (int d = 0; d < n_dev; ++d) devs[d] = d;
    Context multi(devs);
    CHECK(multi.device_count() == n_dev);
    Rng rng(6);
    HashesWithDistanceSet sets(5, 100, 150, 50, rng);
    auto all = sets.all_members(rng);
    name_all(all);
    for (double tol : {0.05, 0.1}) {
        const auto a = search(all, tol, multi);
        const auto b = search(all, tol, ctx);
        CHECK(a.size() == b.size() && !a.empty());
        for (size_t k = 0; k < a.size(); ++k) CHECK(a[k].duplicates() == b[k].duplicates());
    }
    std::vector<VideoHash> refs = {sets.groups[0].start_hash.with_src_path("ref0"), sets.groups[4].start_hash.with_src_path("ref4")};
    const auto ra = search_with_references(refs, all, 0.05, multi);
    const auto rb = search_with_references(refs, all, 0.05, ctx);
    CHECK(ra.size() == rb.size() && ra.size() == 2);
    for (size_t k = 0; k < ra.size(); ++k) CHECK(ra[k].duplicates() == rb[k].duplicates() && ra[k].reference() == rb[k].reference());
    return 0;
}

static int test_sort_order_host_only() {
    std::vector<VideoHash> v;
    const char* pieces[] = {"a", "b", "-", ".", "/", "..", "_", "0", "v", "\x01", " ", "ab", "~"};
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13, s ^= s >> 7, s ^= s << 17; return s; };
    for (int i = 0; i < 40000; ++i) {
        std::string p;
        const int n = 1 + (int)(rnd() % 8);
        for (int k = 0; k < n; ++k) p += pieces[rnd() % 13];
        v.push_back(VideoHash::empty_hash(p).with_duration((uint32_t)(rnd() % 3)));
    }
    for (int i = 0; i < 500; ++i) v.push_back(v[(size_t)i * 7]);  // exact ties: stability
    std::vector<uint64_t> want(v.size());
    for (size_t i = 0; i < want.size(); ++i) want[i] = i;
    std::stable_sort(want.begin(), want.end(), [&](uint64_t a, uint64_t b) {
        if (v[a].duration() != v[b].duration()) return v[a].duration() < v[b].duration();
        return path_cmp(v[a].src_path(), v[b].src_path()) < 0;
    });
    CHECK(detail::sort_order(v) == want);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--host-only") {
        int r = test_sort_order_host_only();
        std::printf("%s test_sort_order_host_only\n", r ? "FAILED" : "ok    ");
        return r;
    }
    try {
        Context ctx(0);
        int fails = 0;
#define RUN(t)                                          \
    do {                                                \
        int r = t(ctx);                                 \
        std::printf("%s %s\n", r ? "FAILED" : "ok    ", #t); \
        fails += r;                                     \
    } while (0)
        RUN(test_metric);
        RUN(test_searching_nothing_returns_empty_vec);
        RUN(test_find_dups_finds_a_known_group);
        RUN(test_find_dups_discriminates_by_duration);
        RUN(test_find_dups_discriminates_by_distance);
        RUN(test_find_with_refs);
        RUN(test_group_order_and_paths);
        RUN(test_hash_frames);
        RUN(test_prepared_search_object);
        RUN(test_multi_device_context);
        std::printf("%s\n", fails ? "SOME TESTS FAILED" : "ALL TESTS PASSED");
        return fails ? 1 : 0;
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
}
