// hvb_encoder.cpp -- the submission queue of include/hvb_encoder.h.
//
// What it replaces in the reference: nothing -- the reference's pool threads call the havoc tables directly
// (turing/Search.hpp:1470, :1978-1982, turing/Reconstruct.cpp:244-353) because a CPU primitive costs less than a
// hand-over.  On the device a launch costs more than a primitive, so the calls of all pool threads that are in flight
// at the same time (up to PicHeightInCtbs x concurrent-frames of them, turing/TaskEncodeSubstream.cpp:55-136) are gathered
// into one batch per kind.  The session is built on the public hvb.h ABI only.
//
// Threading: workers append to the buffer being filled (double-buffered, page-locked, device-addressable) under one
// mutex and sleep on their own condition variable; the dispatcher thread flips the buffers, issues every kind's batch on
// the context's stream, waits for the stream once, hands the results out and wakes the workers.  The hvb_context is
// touched by the dispatcher thread only.
#include "../../include/hvb_encoder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Waiter
{
    std::mutex m;
    std::condition_variable cv;
    bool done = false;
    int rc = 0;
};

Waiter &myWaiter()
{
    static thread_local Waiter w;
    return w;
}

struct Request
{
    Waiter *waiter;
    void *dst;      // where the caller wants the results of its tasks
    int first, count;
};

// one kind of task: two page-locked task / result arrays, the one being filled and the one on the device
template <class Task, class Result, int ResultsPerTask = 1>
struct Lane
{
    int capacity = 0;
    Task *tasks[2] = {nullptr, nullptr};
    Result *results[2] = {nullptr, nullptr};
    std::vector<Request> requests[2];
    int n[2] = {0, 0};
    int64_t totalTasks = 0, totalBatches = 0;

    int init(hvb_context *ctx, int cap)
    {
        capacity = cap;
        for (int b = 0; b < 2; ++b)
        {
            int rc = hvb_host_alloc(ctx, sizeof(Task) * cap, reinterpret_cast<void **>(&tasks[b]));
            if (rc) return rc;
            rc = hvb_host_alloc(ctx, sizeof(Result) * ResultsPerTask * cap, reinterpret_cast<void **>(&results[b]));
            if (rc) return rc;
        }
        return 0;
    }
    void release(hvb_context *ctx)
    {
        for (int b = 0; b < 2; ++b)
        {
            if (tasks[b]) hvb_host_free(ctx, tasks[b]);
            if (results[b]) hvb_host_free(ctx, results[b]);
        }
    }
    // results of buffer b back to the callers; the waiters are collected for one wake-up pass
    void deliver(int b, int rc, std::vector<Waiter *> &wake)
    {
        for (const Request &r : requests[b])
        {
            if (r.dst && !rc) memcpy(r.dst, results[b] + (size_t)r.first * ResultsPerTask, sizeof(Result) * ResultsPerTask * r.count);
            r.waiter->rc = rc;
            wake.push_back(r.waiter);
        }
        totalTasks += n[b];
        totalBatches += n[b] ? 1 : 0;
        requests[b].clear();
        n[b] = 0;
    }
};

struct UploadTask
{
    int pic, cIdx, x0, y0, w, h;
    size_t offset; // into the staging buffer, bytes
};

struct TuExtra // per transform block: where the caller wants reconstruction and levels
{
    void *rec;
    intptr_t recStride;
    int16_t *levels;
    int cell, log2n;
};

constexpr int kTuCell = 32;          // staging cell of a transform block: 32 x 32 samples
constexpr int kTuCellsPerRow = 16;   // wrapped staging pictures are 512 samples wide
constexpr int kTuCapacity = 2048;    // transform blocks per batch
constexpr int kRdoqSnapshots = 512;  // context snapshots per batch
constexpr size_t kUploadBytes = 24u << 20;
constexpr size_t kPoolSamples = 1u << 20; // intra neighbours per batch

} // namespace

struct hvbenc
{
    hvb_context *ctx = nullptr;
    int bps = 1, bitDepth = 8, width = 0, height = 0;
    std::string lastError;

    std::mutex m;
    std::condition_variable workCv, spaceCv;
    int fill = 0;         // buffer the workers append to
    int pending = 0;      // requests in the buffer being filled
    bool stop = false;
    std::thread dispatcher;

    Lane<hvb_me_task, hvb_me_result> me;
    Lane<hvb_me_bi_task, hvb_me_bi_result> bi;
    Lane<hvb_pu_cost_task, int32_t, 3> pu;
    Lane<hvb_intra_sweep_task, int32_t, 35> intra;
    Lane<hvb_tu_task, hvb_tu_result> tu;

    // uploads: rectangles copied into a page-locked staging buffer by the caller
    std::vector<UploadTask> uploads[2];
    std::vector<Waiter *> uploadWaiters[2];
    char *uploadStage[2] = {nullptr, nullptr};
    size_t uploadUsed[2] = {0, 0};
    int64_t totalUploads = 0, totalUploadBytes = 0;

    // intra neighbours of the batch being filled (samples)
    char *poolStage[2] = {nullptr, nullptr};
    size_t poolUsed[2] = {0, 0};

    // transform blocks: prediction / reconstruction cells in two wrapped pictures per buffer, snapshots, levels
    char *tuPredHost[2] = {nullptr, nullptr}, *tuRecHost[2] = {nullptr, nullptr};
    int tuPredPic[2] = {-1, -1}, tuRecPic[2] = {-1, -1};
    std::vector<TuExtra> tuExtra[2];
    hvb_rdoq_ctx *snapshots[2] = {nullptr, nullptr};
    int nSnapshots[2] = {0, 0};
    int16_t *levelsHost = nullptr; // page-locked, kTuCapacity * 1024

    // picture pool
    struct Slot
    {
        const void *key = nullptr;
        uint64_t lastUse = 0;
        int pic = -1;
    };
    std::vector<Slot> slots;
    std::unordered_map<const void *, int> byKey;
    uint64_t useClock = 0;

    double deviceSeconds = 0;
    int64_t dispatches = 0;
};

namespace {

int fail(hvbenc *enc, int rc, const char *what)
{
    if (enc) enc->lastError = std::string(what) + ": " + (enc->ctx ? hvb_last_error(enc->ctx) : "");
    return rc;
}

// Issue everything in buffer b.  Order on the stream: uploads, then the searches, costs, sweeps and transform blocks.
int runBatch(hvbenc *enc, int b)
{
    hvb_context *ctx = enc->ctx;
    int rc = 0;
    for (const UploadTask &u : enc->uploads[b])
    {
        rc = hvb_picture_upload_rect(ctx, u.pic, u.cIdx, enc->uploadStage[b] + u.offset, u.w, u.x0, u.y0, u.w, u.h);
        if (rc) return rc;
    }
    if (enc->me.n[b]) rc = hvb_me_search_batch(ctx, enc->me.tasks[b], enc->me.n[b], enc->me.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (enc->bi.n[b]) rc = hvb_me_bi_search_batch(ctx, enc->bi.tasks[b], enc->bi.n[b], enc->bi.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (enc->pu.n[b]) rc = hvb_pu_cost_batch(ctx, enc->pu.tasks[b], enc->pu.n[b], enc->pu.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (enc->intra.n[b])
    {
        rc = hvb_pool_upload(ctx, enc->poolStage[b], enc->poolUsed[b], 0);
        if (!rc) rc = hvb_intra_satd35_batch(ctx, enc->intra.tasks[b], enc->intra.n[b], enc->intra.results[b], HVB_DEVICE);
        if (rc) return rc;
    }
    size_t levelCount = 0;
    if (enc->tu.n[b])
    {
        if (enc->nSnapshots[b]) rc = hvb_rdoq_contexts_upload(ctx, enc->snapshots[b], enc->nSnapshots[b], 0);
        if (!rc) rc = hvb_tu_chain_batch(ctx, enc->tu.tasks[b], enc->tu.n[b], enc->tu.results[b], HVB_DEVICE);
        if (rc) return rc;
        levelCount = (size_t)enc->tu.n[b] * 1024;
    }
    // one wait for the whole batch; the levels come back with it
    if (levelCount) rc = hvb_coeff_download(ctx, enc->levelsHost, levelCount, 0);
    else rc = hvb_sync(ctx);
    if (rc) return rc;
    // reconstruction cells and levels back to the callers' buffers
    for (size_t i = 0; i < enc->tuExtra[b].size(); ++i)
    {
        const TuExtra &x = enc->tuExtra[b][i];
        const int n = 1 << x.log2n;
        const size_t cellStride = (size_t)kTuCell * kTuCellsPerRow * enc->bps;
        const char *cell = enc->tuRecHost[b] + (size_t)(x.cell / kTuCellsPerRow) * kTuCell * cellStride +
                           (size_t)(x.cell % kTuCellsPerRow) * kTuCell * enc->bps;
        for (int y = 0; y < n; ++y)
            memcpy(static_cast<char *>(x.rec) + (size_t)y * x.recStride * enc->bps, cell + y * cellStride, (size_t)n * enc->bps);
        memcpy(x.levels, enc->levelsHost + i * 1024, sizeof(int16_t) * n * n);
    }
    return 0;
}

void dispatch(hvbenc *enc)
{
    std::vector<Waiter *> wake;
    for (;;)
    {
        int b;
        {
            std::unique_lock<std::mutex> lock(enc->m);
            enc->workCv.wait(lock, [&] { return enc->pending > 0 || enc->stop; });
            if (!enc->pending && enc->stop) return;
            b = enc->fill;
            enc->fill ^= 1;
            enc->pending = 0;
        }
        enc->spaceCv.notify_all();
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = runBatch(enc, b);
        enc->deviceSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ++enc->dispatches;
        wake.clear();
        enc->me.deliver(b, rc, wake);
        enc->bi.deliver(b, rc, wake);
        enc->pu.deliver(b, rc, wake);
        enc->intra.deliver(b, rc, wake);
        enc->tu.deliver(b, rc, wake);
        for (Waiter *w : enc->uploadWaiters[b])
        {
            w->rc = rc;
            wake.push_back(w);
        }
        enc->totalUploads += (int64_t)enc->uploads[b].size();
        enc->totalUploadBytes += (int64_t)enc->uploadUsed[b];
        enc->uploads[b].clear();
        enc->uploadWaiters[b].clear();
        enc->uploadUsed[b] = 0;
        enc->poolUsed[b] = 0;
        enc->tuExtra[b].clear();
        enc->nSnapshots[b] = 0;
        if (rc) enc->lastError = std::string("batch failed: ") + hvb_last_error(enc->ctx);
        // a worker may have several requests in one batch (never the case today); wake each once
        std::sort(wake.begin(), wake.end());
        wake.erase(std::unique(wake.begin(), wake.end()), wake.end());
        for (Waiter *w : wake)
        {
            {
                std::lock_guard<std::mutex> g(w->m);
                w->done = true;
            }
            w->cv.notify_one();
        }
    }
}

int waitFor(Waiter &w)
{
    std::unique_lock<std::mutex> lock(w.m);
    w.cv.wait(lock, [&] { return w.done; });
    return w.rc;
}

// append `count` tasks of a lane; `extra(b, first)` runs under the lock once the room is there
template <class LaneT, class Task, class Extra>
int submit(hvbenc *enc, LaneT &lane, const Task *tasks, int count, void *dst, Extra extra)
{
    if (!enc || !tasks || count <= 0 || count > lane.capacity) return HVB_ERR_INVALID;
    Waiter &w = myWaiter();
    w.done = false;
    {
        std::unique_lock<std::mutex> lock(enc->m);
        enc->spaceCv.wait(lock, [&] { return lane.n[enc->fill] + count <= lane.capacity && extra(enc->fill, -1); });
        const int b = enc->fill, first = lane.n[b];
        memcpy(lane.tasks[b] + first, tasks, sizeof(Task) * count);
        extra(b, first);
        lane.requests[b].push_back(Request{&w, dst, first, count});
        lane.n[b] += count;
        ++enc->pending;
    }
    enc->workCv.notify_one();
    return waitFor(w);
}

} // namespace

extern "C" int hvbenc_create(int device, int bytes_per_sample, int bit_depth, int width, int height, int pool_pictures, hvbenc **out)
{
    if (!out || pool_pictures < 2 || pool_pictures > 200) return HVB_ERR_INVALID;
    *out = nullptr;
    hvb_context *ctx = nullptr;
    int rc = hvb_create(device, bytes_per_sample, bit_depth, &ctx);
    if (rc) return rc;
    hvbenc *enc = new hvbenc;
    enc->ctx = ctx;
    enc->bps = bytes_per_sample;
    enc->bitDepth = bit_depth;
    enc->width = width;
    enc->height = height;
    rc = hvb_set_pipelined(ctx, 1); // uploads from the page-locked staging buffers are enqueued, not waited for
    if (!rc) rc = enc->me.init(ctx, 4096);
    if (!rc) rc = enc->bi.init(ctx, 4096);
    if (!rc) rc = enc->pu.init(ctx, 8192);
    if (!rc) rc = enc->intra.init(ctx, 4096);
    if (!rc) rc = enc->tu.init(ctx, kTuCapacity);
    const size_t cellBytes = (size_t)kTuCell * kTuCellsPerRow * bytes_per_sample * kTuCell * (kTuCapacity / kTuCellsPerRow);
    for (int b = 0; b < 2 && !rc; ++b)
    {
        rc = hvb_host_alloc(ctx, kUploadBytes, reinterpret_cast<void **>(&enc->uploadStage[b]));
        if (!rc) rc = hvb_host_alloc(ctx, kPoolSamples * bytes_per_sample, reinterpret_cast<void **>(&enc->poolStage[b]));
        if (!rc) rc = hvb_host_alloc(ctx, cellBytes, reinterpret_cast<void **>(&enc->tuPredHost[b]));
        if (!rc) rc = hvb_host_alloc(ctx, cellBytes, reinterpret_cast<void **>(&enc->tuRecHost[b]));
        if (!rc) rc = hvb_host_alloc(ctx, sizeof(hvb_rdoq_ctx) * kRdoqSnapshots, reinterpret_cast<void **>(&enc->snapshots[b]));
        if (!rc) rc = hvb_picture_wrap(ctx, enc->tuPredHost[b], kTuCell * kTuCellsPerRow, kTuCell * kTuCellsPerRow, kTuCell * (kTuCapacity / kTuCellsPerRow), &enc->tuPredPic[b]);
        if (!rc) rc = hvb_picture_wrap(ctx, enc->tuRecHost[b], kTuCell * kTuCellsPerRow, kTuCell * kTuCellsPerRow, kTuCell * (kTuCapacity / kTuCellsPerRow), &enc->tuRecPic[b]);
    }
    if (!rc) rc = hvb_host_alloc(ctx, sizeof(int16_t) * 1024 * kTuCapacity, reinterpret_cast<void **>(&enc->levelsHost));
    if (!rc)
    {
        // size the level pool once: block i of a batch owns elements [1024 i, 1024 (i + 1))
        std::vector<int16_t> zero(1024, 0);
        rc = hvb_coeff_upload(ctx, zero.data(), 1024, (size_t)1024 * (kTuCapacity - 1));
    }
    enc->slots.resize(pool_pictures);
    for (int i = 0; i < pool_pictures && !rc; ++i) rc = hvb_picture_create(ctx, width, height, 96, &enc->slots[i].pic);
    if (!rc) rc = hvb_sync(ctx);
    if (rc)
    {
        fprintf(stderr, "hvbenc_create: %s\n", hvb_last_error(ctx));
        hvb_destroy(ctx);
        delete enc;
        return rc;
    }
    enc->dispatcher = std::thread(dispatch, enc);
    *out = enc;
    return HVB_OK;
}

extern "C" void hvbenc_destroy(hvbenc *enc)
{
    if (!enc) return;
    {
        std::lock_guard<std::mutex> g(enc->m);
        enc->stop = true;
    }
    enc->workCv.notify_all();
    if (enc->dispatcher.joinable()) enc->dispatcher.join();
    hvb_sync(enc->ctx);
    enc->me.release(enc->ctx);
    enc->bi.release(enc->ctx);
    enc->pu.release(enc->ctx);
    enc->intra.release(enc->ctx);
    enc->tu.release(enc->ctx);
    for (int b = 0; b < 2; ++b)
    {
        if (enc->uploadStage[b]) hvb_host_free(enc->ctx, enc->uploadStage[b]);
        if (enc->poolStage[b]) hvb_host_free(enc->ctx, enc->poolStage[b]);
        if (enc->tuPredHost[b]) hvb_host_free(enc->ctx, enc->tuPredHost[b]);
        if (enc->tuRecHost[b]) hvb_host_free(enc->ctx, enc->tuRecHost[b]);
        if (enc->snapshots[b]) hvb_host_free(enc->ctx, enc->snapshots[b]);
    }
    if (enc->levelsHost) hvb_host_free(enc->ctx, enc->levelsHost);
    hvb_destroy(enc->ctx);
    delete enc;
}

extern "C" const char *hvbenc_last_error(hvbenc *enc) { return enc ? enc->lastError.c_str() : "null session"; }

extern "C" int hvbenc_picture(hvbenc *enc, const void *key, int fresh, int *pic)
{
    if (!enc || !key || !pic) return HVB_ERR_INVALID;
    std::lock_guard<std::mutex> g(enc->m);
    auto it = enc->byKey.find(key);
    int slot;
    if (it != enc->byKey.end())
        slot = it->second;
    else
    {
        // least recently used slot (unbound slots have lastUse 0)
        slot = 0;
        for (size_t i = 1; i < enc->slots.size(); ++i)
            if (enc->slots[i].lastUse < enc->slots[slot].lastUse) slot = (int)i;
        if (enc->slots[slot].key) enc->byKey.erase(enc->slots[slot].key);
        enc->slots[slot].key = key;
        enc->byKey[key] = slot;
    }
    (void)fresh; // a renewed binding keeps its slot: every sample a search may read is uploaded before the search is issued
    enc->slots[slot].lastUse = ++enc->useClock;
    *pic = enc->slots[slot].pic;
    return HVB_OK;
}

extern "C" int hvbenc_upload_rect(hvbenc *enc, int pic, int cIdx, const void *host, intptr_t stride, int x0, int y0, int w, int h)
{
    if (!enc || !host || w <= 0 || h <= 0 || stride < w) return HVB_ERR_INVALID;
    const size_t rowBytes = (size_t)w * enc->bps, bytes = (rowBytes * h + 63) & ~size_t(63);
    if (bytes > kUploadBytes)
    {
        // larger than the staging buffer: split by rows
        const int rows = std::max(1, (int)(kUploadBytes / 2 / rowBytes));
        for (int y = 0; y < h; y += rows)
        {
            const int rc = hvbenc_upload_rect(enc, pic, cIdx, static_cast<const char *>(host) + (size_t)y * stride * enc->bps, stride, x0, y0 + y, w,
                                              std::min(rows, h - y));
            if (rc) return rc;
        }
        return HVB_OK;
    }
    Waiter &wt = myWaiter();
    wt.done = false;
    {
        std::unique_lock<std::mutex> lock(enc->m);
        enc->spaceCv.wait(lock, [&] { return enc->uploadUsed[enc->fill] + bytes <= kUploadBytes; });
        const int b = enc->fill;
        char *dst = enc->uploadStage[b] + enc->uploadUsed[b];
        for (int y = 0; y < h; ++y) memcpy(dst + y * rowBytes, static_cast<const char *>(host) + (size_t)y * stride * enc->bps, rowBytes);
        enc->uploads[b].push_back(UploadTask{pic, cIdx, x0, y0, w, h, enc->uploadUsed[b]});
        enc->uploadWaiters[b].push_back(&wt);
        enc->uploadUsed[b] += bytes;
        ++enc->pending;
    }
    enc->workCv.notify_one();
    return waitFor(wt);
}

extern "C" int hvbenc_me(hvbenc *enc, const hvb_me_task *task, hvb_me_result *out)
{
    return submit(enc, enc->me, task, 1, out, [](int, int) { return true; });
}

extern "C" int hvbenc_me_bi(hvbenc *enc, const hvb_me_bi_task *task, hvb_me_bi_result *out)
{
    return submit(enc, enc->bi, task, 1, out, [](int, int) { return true; });
}

extern "C" int hvbenc_pu_cost(hvbenc *enc, const hvb_pu_cost_task *tasks, int n, int32_t *out)
{
    return submit(enc, enc->pu, tasks, n, out, [](int, int) { return true; });
}

extern "C" int hvbenc_intra_sweep(hvbenc *enc, const hvb_intra_sweep_task *task, const void *neighbours, int32_t *out)
{
    if (!enc || !task || !neighbours || task->log2n < 2 || task->log2n > 5) return HVB_ERR_INVALID;
    const size_t count = (size_t)(4 << task->log2n) + 1, room = (count + 15) & ~size_t(15);
    return submit(enc, enc->intra, task, 1, out, [&](int b, int first) {
        if (first < 0) return enc->poolUsed[b] + room <= kPoolSamples;
        const size_t at = enc->poolUsed[b];
        memcpy(enc->poolStage[b] + at * enc->bps, neighbours, count * enc->bps);
        hvb_intra_sweep_task &t = enc->intra.tasks[b][first];
        t.nb_unfiltered = (int32_t)(at + (size_t)(2 << task->log2n)); // index of p(-1,-1)
        t.nb_filtered = -1;                                           // derived on the device
        enc->poolUsed[b] += room;
        return true;
    });
}

extern "C" int hvbenc_tu_chain(hvbenc *enc, hvb_tu_task *tasks, int n, const hvb_rdoq_ctx *snapshot, const void *const *pred, const intptr_t *pred_stride,
                               void *const *rec, const intptr_t *rec_stride, int16_t *const *levels, hvb_tu_result *out)
{
    if (!enc || !tasks || n <= 0 || !pred || !pred_stride || !rec || !rec_stride || !levels || !out) return HVB_ERR_INVALID;
    return submit(enc, enc->tu, tasks, n, out, [&](int b, int first) {
        if (first < 0) return !snapshot || enc->nSnapshots[b] < kRdoqSnapshots;
        int snap = 0;
        if (snapshot)
        {
            snap = enc->nSnapshots[b]++;
            enc->snapshots[b][snap] = *snapshot;
        }
        const size_t cellStride = (size_t)kTuCell * kTuCellsPerRow * enc->bps;
        for (int i = 0; i < n; ++i)
        {
            hvb_tu_task &t = enc->tu.tasks[b][first + i];
            const int cell = first + i, nn = 1 << t.log2n;
            const int cx = (cell % kTuCellsPerRow) * kTuCell, cy = (cell / kTuCellsPerRow) * kTuCell;
            char *dst = enc->tuPredHost[b] + (size_t)cy * cellStride + (size_t)cx * enc->bps;
            for (int y = 0; y < nn; ++y)
                memcpy(dst + y * cellStride, static_cast<const char *>(pred[i]) + (size_t)y * pred_stride[i] * enc->bps, (size_t)nn * enc->bps);
            t.pred = hvb_block{(int16_t)enc->tuPredPic[b], 0, (int16_t)cx, (int16_t)cy};
            t.rec = hvb_block{(int16_t)enc->tuRecPic[b], 0, (int16_t)cx, (int16_t)cy};
            t.levels = cell * 1024;
            t.rdoq_ctx = snap;
            enc->tuExtra[b].push_back(TuExtra{rec[i], rec_stride[i], levels[i], cell, t.log2n});
        }
        return true;
    });
}

extern "C" int hvbenc_stats(hvbenc *enc, char *buf, size_t bytes)
{
    if (!enc || !buf || !bytes) return HVB_ERR_INVALID;
    std::lock_guard<std::mutex> g(enc->m);
    snprintf(buf, bytes,
             "{\"dispatches\": %lld, \"device_wait_s\": %.3f, \"kernel_launches\": %lld, "
             "\"me\": {\"tasks\": %lld, \"batches\": %lld}, \"me_bi\": {\"tasks\": %lld, \"batches\": %lld}, "
             "\"pu_cost\": {\"tasks\": %lld, \"batches\": %lld}, \"intra_sweep\": {\"tasks\": %lld, \"batches\": %lld}, "
             "\"tu_chain\": {\"tasks\": %lld, \"batches\": %lld}, \"uploads\": {\"rects\": %lld, \"bytes\": %lld}}",
             (long long)enc->dispatches, enc->deviceSeconds, (long long)hvb_launch_count(enc->ctx), (long long)enc->me.totalTasks,
             (long long)enc->me.totalBatches, (long long)enc->bi.totalTasks, (long long)enc->bi.totalBatches, (long long)enc->pu.totalTasks,
             (long long)enc->pu.totalBatches, (long long)enc->intra.totalTasks, (long long)enc->intra.totalBatches,
             (long long)enc->tu.totalTasks, (long long)enc->tu.totalBatches, (long long)enc->totalUploads, (long long)enc->totalUploadBytes);
    return HVB_OK;
}
