// pipeline.cu -- batch-oriented caller of the hashing path.
//
// The reference hashes one file per rayon worker, end to end: decode, crop, resize, DCT, insert
// (vid_dup_finder_app/src/video_hash_filesystem_cache/video_hash_filesystem_cache.rs:237-257 -> generic_cache_if.rs:25-44
// -> VideoHashBuilder::hash, video_hash_builder.rs:80-83,214-223).  The GPU only pays off on batches, so the loop becomes:
//   decode threads (the caller's, any number)  --vdf_pipeline_push-->  pinned batch buffers (two, filled alternately)
//   one worker thread (owns the context)       --vdf_hash_stacks-->    results queue  --vdf_pipeline_poll--> collector
// Producers copy their decoded frames straight into pinned memory (no second staging copy); while the worker uploads and
// hashes batch b, producers fill batch b+1, so decode, PCIe and kernels overlap.  Results carry the caller's tag, the
// per-stack status (Error::NotEnoughFrames / Error::VidProc as in vdf_hash_stacks), the 16 hash words and the crop.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <thread>

#include "common.cuh"

struct vdf_hash_pipeline {
    struct Slot {
        uint64_t tag;
        vdf_stack_desc desc;
    };
    struct Batch {
        vdf::PinnedBuf buf;
        size_t used = 0;
        std::vector<Slot> slots;
        int writers = 0;     // producers still copying into their reserved region
        bool ready = false;  // closed: waiting for / being processed by the worker
    };
    vdf_ctx* ctx = nullptr;
    int cropdetect = VDF_CROPDETECT_LETTERBOX;
    size_t batch_bytes = 0;
    uint32_t batch_stacks = 0;
    std::mutex m;
    std::condition_variable cv_space, cv_work, cv_results;
    Batch batch[2];
    int fill = 0;  // producers fill batch[fill]
    bool stop = false;
    std::deque<vdf_pipeline_result> results;
    uint64_t pushed = 0, completed = 0;
    int err_code = VDF_OK;
    std::string err;
    std::thread worker;

    void close_fill_locked() {  // hand the batch being filled to the worker, continue in the other one when it is free
        if (batch[fill].slots.empty()) return;
        batch[fill].ready = true;
        cv_work.notify_all();
        fill ^= 1;
    }

    void run() {
        std::vector<vdf_stack_desc> descs;
        std::vector<uint64_t> hashes;
        std::vector<int32_t> status;
        std::vector<uint32_t> crop;
        int next = 0;  // batches are closed alternately, so they are processed alternately
        for (;;) {
            std::unique_lock<std::mutex> lk(m);
            cv_work.wait(lk, [&] { return stop || (batch[next].ready && batch[next].writers == 0); });
            if (!(batch[next].ready && batch[next].writers == 0)) return;  // stop with nothing to do
            Batch& b = batch[next];
            const size_t n = b.slots.size();
            lk.unlock();
            descs.resize(n), hashes.assign(n * 16, 0), status.assign(n, 0), crop.assign(n * 4, 0);
            for (size_t k = 0; k < n; ++k) descs[k] = b.slots[k].desc;
            const int rc = vdf_hash_stacks(ctx, b.buf.as<uint8_t>(), descs.data(), (uint32_t)n, cropdetect, hashes.data(), status.data(),
                                           crop.data());
            lk.lock();
            if (rc != VDF_OK && err_code == VDF_OK) err_code = rc, err = vdf_last_error(ctx);
            for (size_t k = 0; k < n; ++k) {
                vdf_pipeline_result r;
                r.tag = b.slots[k].tag;
                r.status = rc == VDF_OK ? status[k] : rc;
                memcpy(r.hash, &hashes[k * 16], 128);
                memcpy(r.crop, &crop[k * 4], 16);
                results.push_back(r);
            }
            completed += n;
            b.slots.clear(), b.used = 0, b.ready = false;
            next ^= 1;
            cv_space.notify_all();
            cv_results.notify_all();
        }
    }
};

extern "C" {

int vdf_pipeline_create(vdf_ctx* ctx, uint32_t max_batch_stacks, uint64_t batch_bytes, int cropdetect, vdf_hash_pipeline** out) {
    if (!ctx || !out || max_batch_stacks == 0 || batch_bytes == 0) return VDF_ERR_INVALID;
    if (cropdetect != VDF_CROPDETECT_NONE && cropdetect != VDF_CROPDETECT_LETTERBOX) return VDF_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return VDF_ERR_CUDA;
    auto* p = new (std::nothrow) vdf_hash_pipeline;
    if (!p) return VDF_ERR_ALLOC;
    p->ctx = ctx, p->cropdetect = cropdetect, p->batch_bytes = batch_bytes, p->batch_stacks = max_batch_stacks;
    for (auto& b : p->batch) {
        if (b.buf.ensure(batch_bytes) != cudaSuccess) {
            cudaGetLastError();
            p->batch[0].buf.release(), p->batch[1].buf.release();
            delete p;
            return VDF_ERR_ALLOC;
        }
    }
    p->worker = std::thread([p] { p->run(); });
    *out = p;
    return VDF_OK;
}

int vdf_pipeline_push(vdf_hash_pipeline* p, uint64_t tag, const uint8_t* const* frames, uint32_t n_frames, uint32_t width,
                      uint32_t height, uint32_t pitch, uint32_t flags) {
    if (!p || (n_frames && !frames) || pitch < width) return VDF_ERR_INVALID;
    if (!(flags & VDF_STACK_FLAG_MIXED_SIZES) && n_frames >= VDF_DCT_SIZE && (width == 0 || height == 0)) return VDF_ERR_INVALID;
    // only what vdf_hash_stacks reads is staged: the first 16 frames of a well-formed stack (a short or mixed-size
    // stack produces its error status without touching pixels)
    const bool hashed = !(flags & VDF_STACK_FLAG_MIXED_SIZES) && n_frames >= VDF_DCT_SIZE && width && height;
    const uint32_t staged = hashed ? VDF_DCT_SIZE : 0;
    const size_t frame_bytes = (size_t)width * height, bytes = (frame_bytes * staged + 255) & ~(size_t)255;
    if (bytes > p->batch_bytes) return VDF_ERR_INVALID;
    std::unique_lock<std::mutex> lk(p->m);
    for (;;) {
        if (p->stop) return VDF_ERR_INVALID;
        vdf_hash_pipeline::Batch& b = p->batch[p->fill];
        if (!b.ready && b.used + bytes <= p->batch_bytes && b.slots.size() < p->batch_stacks) break;
        if (!b.ready) p->close_fill_locked();          // full: hand it over and try the other one
        else p->cv_space.wait(lk);                     // both closed: wait for the worker
    }
    vdf_hash_pipeline::Batch& b = p->batch[p->fill];
    const size_t off = b.used;
    b.used += bytes;
    vdf_hash_pipeline::Slot s;
    s.tag = tag;
    s.desc = vdf_stack_desc{off, frame_bytes, width, height, width, n_frames, flags, 0};
    b.slots.push_back(s);
    b.writers++;
    p->pushed++;
    lk.unlock();
    uint8_t* dst = b.buf.as<uint8_t>() + off;  // packed: pitch = width
    for (uint32_t f = 0; f < staged; ++f)
        for (uint32_t y = 0; y < height; ++y) memcpy(dst + f * frame_bytes + (size_t)y * width, frames[f] + (size_t)y * pitch, width);
    lk.lock();
    b.writers--;
    if (b.slots.size() >= p->batch_stacks && !b.ready && &b == &p->batch[p->fill]) p->close_fill_locked();
    p->cv_work.notify_all();
    return VDF_OK;
}

int vdf_pipeline_flush(vdf_hash_pipeline* p) {
    if (!p) return VDF_ERR_INVALID;
    std::unique_lock<std::mutex> lk(p->m);
    const uint64_t target = p->pushed;
    if (!p->batch[p->fill].ready) p->close_fill_locked();
    p->cv_results.wait(lk, [&] { return p->completed >= target; });
    return p->err_code;
}

int vdf_pipeline_poll(vdf_hash_pipeline* p, vdf_pipeline_result* out, uint32_t max_results, uint32_t* n_out, int wait) {
    if (!p || !n_out || (max_results && !out)) return VDF_ERR_INVALID;
    std::unique_lock<std::mutex> lk(p->m);
    if (wait) p->cv_results.wait(lk, [&] { return !p->results.empty() || p->completed >= p->pushed; });
    uint32_t n = 0;
    while (n < max_results && !p->results.empty()) {
        out[n++] = p->results.front();
        p->results.pop_front();
    }
    *n_out = n;
    return p->err_code;
}

const char* vdf_pipeline_error(const vdf_hash_pipeline* p) { return p ? p->err.c_str() : "null pipeline"; }

void vdf_pipeline_destroy(vdf_hash_pipeline* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->cv_work.notify_all();
    p->cv_space.notify_all();
    if (p->worker.joinable()) p->worker.join();
    cudaSetDevice(p->ctx->device);
    p->batch[0].buf.release(), p->batch[1].buf.release();
    delete p;
}

}  // extern "C"
