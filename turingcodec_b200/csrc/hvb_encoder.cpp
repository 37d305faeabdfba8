// hvb_encoder.cpp -- the submission queue of include/hvb_encoder.h.
//
// What it replaces in the reference: nothing -- the reference's pool threads call the havoc tables directly
// (turing/Search.hpp:1470, :1978-1982, turing/Reconstruct.cpp:244-353) because a CPU primitive costs less than a
// hand-over.  On the device a launch costs more than a primitive, so the calls of all pool threads that are in flight
// at the same time (up to PicHeightInCtbs x concurrent-frames of them, turing/TaskEncodeSubstream.cpp:55-136) are gathered
// into one batch per kind.  The session is built on the public hvb.h ABI only.
//
// Threading: a session runs several ENGINES, each an hvb_context with its own stream, its own page-locked buffers and its
// own dispatcher thread; the pictures live in engine 0's context and are imported into the others (same ids everywhere).
// A worker hands its request to the engine with the least work in flight: while the system is lightly loaded a request
// starts at once on an idle engine (the kernels of different engines overlap on the device), under load every engine
// batches what arrives while its previous batch is on the device.  Inside an engine workers append to the buffer being
// filled (double-buffered, device-addressable) under the engine's mutex and sleep on their own condition variable; the
// dispatcher flips the buffers, issues every kind's batch on the stream, waits for the stream once, hands the results
// out and wakes the workers.  An hvb_context is touched by its dispatcher thread only.
#include "../../include/hvb_encoder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <time.h>
#ifdef __linux__
#include <sys/prctl.h>
#endif

namespace {

// A caller waiting for its answer.  Blocking callers (a thread per caller) sleep on the condition variable of their thread's
// waiter; cooperative callers (hvbenc_set_thread_hooks: fibers multiplexed on a scheduler thread) have a waiter on their own
// stack, park through the thread's hook and are announced to the scheduler by `notify` -- no futex on either side.
struct Waiter
{
    std::mutex m;
    std::condition_variable cv;
    std::atomic<int> done{0};
    int rc = 0;
    void (*notify)(void *) = nullptr;
    void *notifyArg = nullptr;
    int64_t tSubmit = 0, tFlip = 0, tDone = 0; // ns: handed over, its batch taken by the dispatcher, answered
};

inline int64_t nowNs()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Waiter &myWaiter()
{
    static thread_local Waiter w;
    return w;
}

thread_local hvbenc_thread_hooks tlHooks = {nullptr, nullptr, nullptr};

// the answer is there: last access to a cooperative waiter is the store (it lives on a stack that may unwind right after).
// Cooperative waiters of one scheduler are announced once per batch: `told` collects the schedulers already notified.
void complete(Waiter *w, std::vector<void *> &told)
{
    w->tDone = nowNs();
    if (w->notify)
    {
        void (*const fn)(void *) = w->notify;
        void *const arg = w->notifyArg;
        w->done.store(1, std::memory_order_release);
        if (std::find(told.begin(), told.end(), arg) == told.end())
        {
            told.push_back(arg);
            fn(arg);
        }
        return;
    }
    {
        std::lock_guard<std::mutex> g(w->m);
        w->done.store(1, std::memory_order_release);
    }
    w->cv.notify_one();
}

struct Request
{
    Waiter *waiter;
    void *dst;      // where the caller wants the results of its tasks
    int first, count;
};

// One page-locked allocation per engine, carved: pinning memory is slow, and an engine needs some thirty buffers.
struct Arena
{
    hvb_context *ctx = nullptr;
    char *base = nullptr;
    size_t size = 0, used = 0;
    static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
    int reserve(hvb_context *c, size_t bytes)
    {
        ctx = c;
        size = bytes;
        used = 0;
        return hvb_host_alloc(c, bytes, reinterpret_cast<void **>(&base));
    }
    template <class T> int take(T **out, size_t bytes)
    {
        *out = nullptr;
        if (!bytes) return 0;
        if (!base || used + padded(bytes) > size) return HVB_ERR_NOMEM;
        *out = reinterpret_cast<T *>(base + used);
        used += padded(bytes);
        return 0;
    }
    void release()
    {
        if (base) hvb_host_free(ctx, base);
        base = nullptr;
    }
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

    static size_t bytes(int cap) { return 2 * (Arena::padded(sizeof(Task) * cap) + Arena::padded(sizeof(Result) * ResultsPerTask * cap)); }
    int init(Arena &arena, int cap)
    {
        capacity = cap;
        for (int b = 0; b < 2; ++b)
        {
            int rc = arena.take(&tasks[b], sizeof(Task) * cap);
            if (rc) return rc;
            rc = arena.take(&results[b], sizeof(Result) * ResultsPerTask * cap);
            if (rc) return rc;
        }
        return 0;
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
constexpr int kTuCapacity = 1024;    // transform blocks per batch
constexpr int kRdoqSnapshots = 256;  // context snapshots per batch
constexpr size_t kUploadBytes = 8u << 20;
constexpr size_t kPoolSamples = 1u << 18; // intra neighbours per batch

} // namespace

struct Engine
{
    hvb_context *ctx = nullptr;
    int bps = 1, bitDepth = 8, width = 0, height = 0;
    std::string lastError;
    std::atomic<int> inflight{0}; // requests handed to this engine and not yet answered
    struct hvbenc *session = nullptr;
    // completion of the batch on the device, signalled by the session's poller thread
    std::mutex doneM;
    std::condition_variable doneCv;
    int pollResult = 0; // 0: in flight, 1: complete, < 0: error
    // the batch's last stream operation stores `sequence` here (page-locked): the completion thread reads memory, not the driver
    int32_t *doneFlag = nullptr;
    int32_t sequence = 0;
    std::chrono::steady_clock::time_point issuedAt;

    std::mutex m;
    std::condition_variable workCv, spaceCv;
    int fill = 0;         // buffer the workers append to
    std::atomic<int> pending{0}; // requests in the buffer being filled (written under m; the service threads peek at it)
    int flying = -1;      // service mode: the buffer whose batch is on the device
    std::chrono::steady_clock::time_point flyingSince;
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
    Arena arena;                   // every page-locked buffer of the engine
    bool serves[6] = {true, true, true, true, true, true}; // kinds this engine can be handed (engines by kind: its own group's only)

    double deviceSeconds = 0;
    int64_t dispatches = 0;
    // statistics: device time per kind (HVB_PROFILE), the kernels' algorithmic bytes, bytes crossing the bus
    bool profile = false;
    double kindMs[6] = {0, 0, 0, 0, 0, 0}; // uploads, me, bi, pu, intra, tu
    double algBytes[6] = {0, 0, 0, 0, 0, 0};
    double toDevice = 0, fromDevice = 0;
};

struct hvbenc
{
    std::vector<Engine *> engines;
    int bps = 1;
    std::string lastError;
    std::mutex poolMutex;
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

    // One completion thread for all engines: it polls the contexts that have a batch in flight (hvb_poll) and wakes the
    // engine whose batch has finished, so that no dispatcher parks a core inside the driver's stream wait.
    bool usePoller = true;
    std::thread poller;
    std::mutex pollM;
    std::condition_variable pollCv;
    std::vector<Engine *> watching;
    bool pollStop = false;
    int pollSleepUs = 20; // HVB_POLL_SLEEP_US; 0: spin
    // Service mode (HVB_SERVICE_THREADS > 0; off by default): no dispatcher thread per engine and no completion thread.  A few
    // service threads sweep the engines: a batch whose flag has arrived is answered and, in the same visit, the requests that
    // gathered meanwhile are issued.  Between a batch's completion and its callers' wake-up there is no thread hand-over left.
    std::vector<std::thread> services;
    std::atomic<bool> serviceStop{false};

    // per kind: requests and the time from hand-over to wake-up (ns), for the statistics
    std::atomic<int64_t> waitNs[6], waitCount[6];
    // where a hand-over's time goes: [0] until the dispatcher takes its batch, [1] the batch (issue, device, completion seen),
    // [2] from the answer to the caller running again
    std::atomic<int64_t> phaseNs[6][3];

    // Engines by kind (0 uploads, 1 me, 2 bi, 3 pu cost, 4 intra sweep, 5 transform blocks).  A batch is issued on one
    // stream in a fixed order, so a sweep (tens of microseconds on the device) that shares a batch with a motion search waits
    // for the search; with enough engines (HVB_ENGINES >= 12) every kind gets its own group of engines -- first[k] .. first[k+1]
    // -- in proportion to its share of the device time of a 4K medium encode (HVB_ENGINE_SHARES overrides), and kernels of
    // different kinds overlap on the device.  With fewer engines every engine serves every kind.
    int first[7] = {0, 0, 0, 0, 0, 0, 0};
    bool partitioned = false;

    void partition()
    {
        const int n = (int)engines.size();
        double share[6] = {2, 6, 7, 3, 5, 9};
        if (const char *v = getenv("HVB_ENGINE_SHARES")) sscanf(v, "%lf,%lf,%lf,%lf,%lf,%lf", &share[0], &share[1], &share[2], &share[3], &share[4], &share[5]);
        partitioned = n >= 12;
        if (const char *v = getenv("HVB_ENGINE_PARTITION")) partitioned = atoi(v) != 0 && n >= 6;
        if (!partitioned) return;
        double total = 0, acc = 0;
        for (double v : share) total += v;
        int count[6], used = 0;
        for (int k = 0; k < 6; ++k) count[k] = 1, ++used;
        // largest remainder over what is left after one engine each
        double want[6];
        for (int k = 0; k < 6; ++k) want[k] = share[k] / total * n - 1;
        while (used < n)
        {
            int best = 0;
            for (int k = 1; k < 6; ++k)
                if (want[k] - (count[k] - 1) > want[best] - (count[best] - 1)) best = k;
            ++count[best], ++used;
        }
        (void)acc;
        for (int k = 0; k < 6; ++k) first[k + 1] = first[k] + count[k];
    }

    Engine *pick(int kind)
    {
        size_t lo = 0, hi = engines.size();
        if (partitioned) lo = (size_t)first[kind], hi = (size_t)first[kind + 1];
        Engine *best = engines[lo];
        int load = best->inflight.load(std::memory_order_relaxed);
        for (size_t i = lo + 1; i < hi && load > 0; ++i)
        {
            const int l = engines[i]->inflight.load(std::memory_order_relaxed);
            if (l < load) best = engines[i], load = l;
        }
        return best;
    }
};

namespace {

int fail(Engine *enc, int rc, const char *what)
{
    if (enc) enc->lastError = std::string(what) + ": " + (enc->ctx ? hvb_last_error(enc->ctx) : "");
    return rc;
}

// Issue everything in buffer b.  Order on the stream: uploads, then the searches, costs, sweeps and transform blocks.
// everything of buffer b onto the engine's stream; does not wait
int issueBatch(Engine *enc, int b)
{
    hvb_context *ctx = enc->ctx;
    int rc = 0;
    const bool prof = enc->profile;
    if (prof) hvb_mark(ctx, 0);
    for (const UploadTask &u : enc->uploads[b])
    {
        rc = hvb_picture_upload_rect(ctx, u.pic, u.cIdx, enc->uploadStage[b] + u.offset, u.w, u.x0, u.y0, u.w, u.h);
        if (rc) return rc;
    }
    if (prof) hvb_mark(ctx, 1);
    if (enc->me.n[b]) rc = hvb_me_search_batch(ctx, enc->me.tasks[b], enc->me.n[b], enc->me.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (prof) hvb_mark(ctx, 2);
    if (enc->bi.n[b]) rc = hvb_me_bi_search_batch(ctx, enc->bi.tasks[b], enc->bi.n[b], enc->bi.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (prof) hvb_mark(ctx, 3);
    if (enc->pu.n[b]) rc = hvb_pu_cost_batch(ctx, enc->pu.tasks[b], enc->pu.n[b], enc->pu.results[b], HVB_DEVICE);
    if (rc) return rc;
    if (prof) hvb_mark(ctx, 4);
    if (enc->intra.n[b])
    {
        rc = hvb_pool_upload(ctx, enc->poolStage[b], enc->poolUsed[b], 0);
        if (!rc) rc = hvb_intra_satd35_batch(ctx, enc->intra.tasks[b], enc->intra.n[b], enc->intra.results[b], HVB_DEVICE);
        if (rc) return rc;
    }
    if (prof) hvb_mark(ctx, 5);
    if (enc->tu.n[b])
    {
        // one launch: the kernel reads this batch's snapshots and writes its levels in the page-locked arrays themselves
        // (the level pool was wrapped once, equipEngine); no upload, no table kernels, no download
        rc = hvb_rdoq_contexts_wrap(ctx, enc->snapshots[b], kRdoqSnapshots);
        if (!rc) rc = hvb_tu_chain_batch(ctx, enc->tu.tasks[b], enc->tu.n[b], enc->tu.results[b], HVB_DEVICE);
        if (rc) return rc;
    }
    if (prof) hvb_mark(ctx, 6);
    return 0;
}

// the batch in buffer b has completed on the device: statistics, reconstructions and levels back to the callers' buffers
int collectBatch(Engine *enc, int b)
{
    hvb_context *ctx = enc->ctx;
    int rc = 0;
    const bool prof = enc->profile;
    // everything has completed.  hvb_sync returns at once and surfaces a sticky error; with the flag scheme the batch's completion
    // is already known, so the three stream queries behind it are only spent now and then
    if (!enc->doneFlag || !enc->session->usePoller || (enc->dispatches & 63) == 0)
    {
        rc = hvb_sync(ctx);
        if (rc) return rc;
    }
    if (prof)
        for (int k = 0; k < 6; ++k)
        {
            float ms = 0;
            if (!hvb_elapsed_ms(ctx, k, k + 1, &ms)) enc->kindMs[k] += ms;
        }
    // statistics: what the kernels of this batch had to touch and what crossed the bus
    {
        const double B = enc->bps;
        for (const UploadTask &u : enc->uploads[b]) enc->toDevice += (double)u.w * u.h * B;
        for (int i = 0; i < enc->me.n[b]; ++i)
        {
            const hvb_me_task &t = enc->me.tasks[b][i];
            const double wh = (double)t.w * t.h, sup = (double)(t.w + 7) * (t.h + 7) + wh;
            enc->algBytes[1] += wh * B + enc->me.results[b][i].nSad * wh * B + (t.halfPel ? (t.quarterPel ? 17 : 9) * sup * B : 0);
        }
        for (int i = 0; i < enc->bi.n[b]; ++i)
        {
            const hvb_me_bi_task &t = enc->bi.tasks[b][i];
            const double wh = (double)t.w * t.h, sup = (double)(t.w + 7) * (t.h + 7) + wh;
            enc->algBytes[2] += wh * B + (t.smallWindow ? 9 : 121) * wh * B + (t.halfPel ? (t.quarterPel ? 18 : 9) : 0) * sup * B;
        }
        for (int i = 0; i < enc->pu.n[b]; ++i)
        {
            const hvb_pu_cost_task &t = enc->pu.tasks[b][i];
            const double sup = (double)(t.w + 7) * (t.h + 7) + (double)t.w * t.h;
            enc->algBytes[3] += 1.5 * sup * B * ((t.ref_pic[0] >= 0) + (t.ref_pic[1] >= 0));
        }
        for (int i = 0; i < enc->intra.n[b]; ++i)
        {
            const double n = 1 << enc->intra.tasks[b][i].log2n;
            enc->algBytes[4] += (4 * n + 1) * B + n * n * B + 140;
        }
        for (int i = 0; i < enc->tu.n[b]; ++i)
        {
            const double nn = (double)(1 << (2 * enc->tu.tasks[b][i].log2n));
            enc->algBytes[5] += 3 * nn * B + 2 * nn + 32;
            enc->toDevice += nn * B;
            enc->fromDevice += nn * B + 2 * nn;
        }
        enc->toDevice += enc->me.n[b] * sizeof(hvb_me_task) + enc->bi.n[b] * sizeof(hvb_me_bi_task) + enc->pu.n[b] * sizeof(hvb_pu_cost_task) +
                         enc->intra.n[b] * sizeof(hvb_intra_sweep_task) + enc->tu.n[b] * sizeof(hvb_tu_task) + enc->poolUsed[b] * B +
                         enc->nSnapshots[b] * sizeof(hvb_rdoq_ctx);
        enc->fromDevice += enc->me.n[b] * sizeof(hvb_me_result) + enc->bi.n[b] * sizeof(hvb_me_bi_result) + enc->pu.n[b] * 12.0 +
                           enc->intra.n[b] * 140.0 + enc->tu.n[b] * sizeof(hvb_tu_result);
    }
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

int runBatch(Engine *enc, int b)
{
    hvb_context *ctx = enc->ctx;
    int rc = issueBatch(enc, b);
    if (rc) return rc;
    // one wait for the whole batch: on the session's completion thread, or (HVB_POLLER=0) inside the driver
    if (!enc->session->usePoller)
    {
        rc = hvb_sync(ctx);
        if (rc) return rc;
    }
    else
    {
        hvbenc *session = enc->session;
        if (enc->doneFlag)
        {
            rc = hvb_signal(ctx, enc->doneFlag, ++enc->sequence);
            if (rc) return rc;
            enc->issuedAt = std::chrono::steady_clock::now();
        }
        {
            std::lock_guard<std::mutex> g(enc->doneM);
            enc->pollResult = 0;
        }
        {
            std::lock_guard<std::mutex> g(session->pollM);
            session->watching.push_back(enc);
        }
        session->pollCv.notify_one();
        std::unique_lock<std::mutex> lock(enc->doneM);
        enc->doneCv.wait(lock, [&] { return enc->pollResult != 0; });
        if (enc->pollResult < 0) return hvb_sync(ctx); // collects the error text
    }
    return collectBatch(enc, b);
}

void answerBatch(Engine *enc, int b, int rc, std::vector<Waiter *> &wake, std::vector<void *> &told);

void stampFlip(Engine *enc, int b)
{
    const int64_t flip = nowNs();
    for (const Request &r : enc->me.requests[b]) r.waiter->tFlip = flip;
    for (const Request &r : enc->bi.requests[b]) r.waiter->tFlip = flip;
    for (const Request &r : enc->pu.requests[b]) r.waiter->tFlip = flip;
    for (const Request &r : enc->intra.requests[b]) r.waiter->tFlip = flip;
    for (const Request &r : enc->tu.requests[b]) r.waiter->tFlip = flip;
    for (Waiter *w : enc->uploadWaiters[b]) w->tFlip = flip;
}

void dispatch(Engine *enc)
{
    std::vector<Waiter *> wake;
    std::vector<void *> told;
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
        stampFlip(enc, b);
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = runBatch(enc, b);
        enc->deviceSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        answerBatch(enc, b, rc, wake, told);
    }
}

// results of buffer b to the callers, buffers reset, callers announced
void answerBatch(Engine *enc, int b, int rc, std::vector<Waiter *> &wake, std::vector<void *> &told)
{
    {
        ++enc->dispatches;
        wake.clear();
        const int answered = (int)(enc->me.requests[b].size() + enc->bi.requests[b].size() + enc->pu.requests[b].size() +
                                   enc->intra.requests[b].size() + enc->tu.requests[b].size() + enc->uploadWaiters[b].size());
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
        enc->inflight.fetch_sub(answered, std::memory_order_relaxed);
        told.clear();
        for (Waiter *w : wake) complete(w, told);
    }
}

void serviceLoop(hvbenc *session, int index, int count)
{
#ifdef __linux__
    prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);
#endif
    std::vector<Waiter *> wake;
    std::vector<void *> told;
    std::vector<Engine *> mine;
    for (size_t e = index; e < session->engines.size(); e += count) mine.push_back(session->engines[e]);
    while (!session->serviceStop.load(std::memory_order_acquire))
    {
        bool any = false;
        for (Engine *enc : mine)
        {
            if (enc->flying >= 0)
            {
                bool done = *const_cast<volatile int32_t *>(enc->doneFlag) == enc->sequence;
                int rc = 0;
                if (!done && std::chrono::steady_clock::now() - enc->issuedAt > std::chrono::milliseconds(200))
                {
                    // a batch that does not come back: ask the driver (a fault never stores the flag)
                    const int r = hvb_poll(enc->ctx);
                    if (r < 0) rc = hvb_sync(enc->ctx), done = true;
                    else if (r > 0) done = true;
                    else enc->issuedAt = std::chrono::steady_clock::now();
                }
                if (!done) continue;
                std::atomic_thread_fence(std::memory_order_acquire);
                const int b = enc->flying;
                if (!rc) rc = collectBatch(enc, b);
                enc->deviceSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - enc->flyingSince).count();
                enc->flying = -1;
                answerBatch(enc, b, rc, wake, told);
                any = true;
            }
            if (enc->flying < 0 && enc->pending.load(std::memory_order_acquire) > 0)
            {
                int b;
                {
                    std::lock_guard<std::mutex> lock(enc->m);
                    b = enc->fill;
                    enc->fill ^= 1;
                    enc->pending = 0;
                }
                enc->spaceCv.notify_all();
                stampFlip(enc, b);
                enc->flyingSince = std::chrono::steady_clock::now();
                int rc = issueBatch(enc, b);
                if (!rc) rc = hvb_signal(enc->ctx, enc->doneFlag, ++enc->sequence);
                enc->issuedAt = std::chrono::steady_clock::now();
                if (rc)
                    answerBatch(enc, b, rc, wake, told);
                else
                    enc->flying = b;
                any = true;
            }
        }
        if (!any)
        {
            timespec ts = {0, (session->pollSleepUs > 0 ? session->pollSleepUs : 1) * 500L};
            nanosleep(&ts, nullptr);
        }
    }
}

void pollLoop(hvbenc *session)
{
#ifdef __linux__
    prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0); // 1 us: the default 50 us slack would triple the sleep below
#endif
    std::vector<Engine *> snapshot;
    for (;;)
    {
        {
            std::unique_lock<std::mutex> lock(session->pollM);
            session->pollCv.wait(lock, [&] { return !session->watching.empty() || session->pollStop; });
            if (session->pollStop && session->watching.empty()) return;
            snapshot = session->watching;
        }
        bool any = false;
        for (Engine *e : snapshot)
        {
            int r;
            if (e->doneFlag)
            {
                r = *const_cast<volatile int32_t *>(e->doneFlag) == e->sequence ? 1 : 0;
                // a batch that has not come back after a long time: ask the driver (a fault never stores the flag)
                if (!r && std::chrono::steady_clock::now() - e->issuedAt > std::chrono::milliseconds(200))
                {
                    r = hvb_poll(e->ctx);
                    if (!r) e->issuedAt = std::chrono::steady_clock::now();
                }
                if (r > 0) std::atomic_thread_fence(std::memory_order_acquire);
            }
            else
                r = hvb_poll(e->ctx);
            if (!r) continue;
            any = true;
            {
                std::lock_guard<std::mutex> g(session->pollM);
                auto &w = session->watching;
                w.erase(std::remove(w.begin(), w.end(), e), w.end());
            }
            {
                std::lock_guard<std::mutex> g(e->doneM);
                e->pollResult = r;
            }
            e->doneCv.notify_one();
        }
        // nothing finished in this sweep: give the core away for a moment (a batch is on the device for 50 .. 500 us; the
        // encoder's own threads need the core more than the sweep needs the last 20 us of latency)
        if (any || session->pollSleepUs <= 0)
            std::this_thread::yield();
        else
        {
            timespec ts = {0, session->pollSleepUs * 1000L};
            nanosleep(&ts, nullptr);
        }
    }
}

int waitFor(Waiter &w)
{
    if (w.notify)
    {
        // cooperative: the scheduler runs other callers on this thread until the answer is there
        while (!w.done.load(std::memory_order_acquire)) tlHooks.park(tlHooks.arg, reinterpret_cast<const volatile int *>(&w.done));
        return w.rc;
    }
    std::unique_lock<std::mutex> lock(w.m);
    w.cv.wait(lock, [&] { return w.done.load(std::memory_order_acquire) != 0; });
    return w.rc;
}

// the waiter of this call: the thread's own, or (cooperative callers) `local` on the caller's stack
Waiter &armWaiter(Waiter &local)
{
    Waiter &w = tlHooks.park ? local : myWaiter();
    w.done.store(0, std::memory_order_relaxed);
    w.notify = tlHooks.park ? tlHooks.notify : nullptr;
    w.notifyArg = tlHooks.arg;
    return w;
}

// append `count` tasks of a lane; `extra(b, first)` runs under the lock once the room is there
void notePhases(hvbenc *session, int kind, const Waiter &w)
{
    const int64_t resumed = nowNs();
    session->phaseNs[kind][0] += w.tFlip - w.tSubmit;
    session->phaseNs[kind][1] += w.tDone - w.tFlip;
    session->phaseNs[kind][2] += resumed - w.tDone;
}

template <class LaneT, class Task, class Extra>
int submit(Engine *enc, int kind, LaneT &lane, const Task *tasks, int count, void *dst, Extra extra)
{
    if (!enc || !tasks || count <= 0 || count > lane.capacity) return HVB_ERR_INVALID;
    Waiter local;
    Waiter &w = armWaiter(local);
    w.tSubmit = nowNs();
    enc->inflight.fetch_add(1, std::memory_order_relaxed);
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
    const int rc = waitFor(w);
    notePhases(enc->session, kind, w);
    return rc;
}

} // namespace

namespace {

int createEngine(int device, int bytes_per_sample, int bit_depth, int width, int height, Engine **out)
{
    hvb_context *ctx = nullptr;
    int rc = hvb_create(device, bytes_per_sample, bit_depth, &ctx);
    if (rc) return rc;
    Engine *enc = new Engine;
    enc->ctx = ctx;
    enc->bps = bytes_per_sample;
    enc->bitDepth = bit_depth;
    enc->width = width;
    enc->height = height;
    if (const char *v = getenv("HVB_PROFILE")) enc->profile = atoi(v) != 0;
    *out = enc;
    return hvb_set_pipelined(ctx, 1); // uploads from the page-locked staging buffers are enqueued, not waited for
}

// buffers and staging pictures of an engine; the session's pictures are already in its context (ids 0 .. pool - 1)
int equipEngine(Engine *enc)
{
    hvb_context *ctx = enc->ctx;
    const int rows = kTuCell * (kTuCapacity / kTuCellsPerRow);
    const size_t cellBytes = (size_t)kTuCell * kTuCellsPerRow * enc->bps * rows;
    // buffers of kinds the engine is never handed (engines by kind) are not allocated: uploads need 16 MB of staging, transform
    // blocks 6 MB of cells and levels
    const size_t uploadBytes = enc->serves[0] ? kUploadBytes : 0, poolBytes = enc->serves[4] ? kPoolSamples * enc->bps : 0;
    const size_t tuCells = enc->serves[5] ? cellBytes : 0, levelBytes = enc->serves[5] ? sizeof(int16_t) * 1024 * kTuCapacity : 0;
    const size_t snapshotBytes = enc->serves[5] ? sizeof(hvb_rdoq_ctx) * kRdoqSnapshots : 0;
    const size_t total = decltype(enc->me)::bytes(1024) + decltype(enc->bi)::bytes(1024) + decltype(enc->pu)::bytes(2048) +
                         decltype(enc->intra)::bytes(1024) + decltype(enc->tu)::bytes(kTuCapacity) +
                         2 * (Arena::padded(uploadBytes) + Arena::padded(poolBytes) + 2 * Arena::padded(tuCells) + Arena::padded(snapshotBytes)) +
                         Arena::padded(levelBytes) + 256;
    int rc = enc->arena.reserve(ctx, total);
    if (!rc) rc = enc->me.init(enc->arena, 1024);
    if (!rc) rc = enc->bi.init(enc->arena, 1024);
    if (!rc) rc = enc->pu.init(enc->arena, 2048);
    if (!rc) rc = enc->intra.init(enc->arena, 1024);
    if (!rc) rc = enc->tu.init(enc->arena, kTuCapacity);
    for (int b = 0; b < 2 && !rc; ++b)
    {
        rc = enc->arena.take(&enc->uploadStage[b], uploadBytes);
        if (!rc) rc = enc->arena.take(&enc->poolStage[b], poolBytes);
        if (!rc) rc = enc->arena.take(&enc->tuPredHost[b], tuCells);
        if (!rc) rc = enc->arena.take(&enc->tuRecHost[b], tuCells);
        if (!rc) rc = enc->arena.take(&enc->snapshots[b], snapshotBytes);
        if (!rc && tuCells) rc = hvb_picture_wrap(ctx, enc->tuPredHost[b], kTuCell * kTuCellsPerRow, kTuCell * kTuCellsPerRow, rows, &enc->tuPredPic[b]);
        if (!rc && tuCells) rc = hvb_picture_wrap(ctx, enc->tuRecHost[b], kTuCell * kTuCellsPerRow, kTuCell * kTuCellsPerRow, rows, &enc->tuRecPic[b]);
    }
    if (!rc) rc = enc->arena.take(&enc->levelsHost, levelBytes);
    const char *flagEnv = getenv("HVB_DONE_FLAG");
    if (!rc && !(flagEnv && atoi(flagEnv) == 0))
    {
        rc = enc->arena.take(&enc->doneFlag, 64);
        if (!rc) *enc->doneFlag = 0;
    }
    if (!rc && levelBytes)
    {
        // block i of a batch owns elements [1024 i, 1024 (i + 1)) of the level pool, which is the page-locked array the results
        // are read from; every batch of the queue takes the one-launch form of the chain
        std::vector<int16_t> zero(1024, 0);
        rc = hvb_coeff_upload(ctx, zero.data(), 1024, 0); // (hvb_tu_chain_batch wants a device pool to exist)
        if (!rc) rc = hvb_coeff_pool_wrap(ctx, enc->levelsHost, (size_t)1024 * kTuCapacity);
        if (!rc) rc = hvb_set_tu_fused_max(ctx, kTuCapacity);
    }
    if (!rc) rc = hvb_sync(ctx);
    return rc;
}

void destroyEngine(Engine *enc)
{
    {
        std::lock_guard<std::mutex> g(enc->m);
        enc->stop = true;
    }
    enc->workCv.notify_all();
    if (enc->dispatcher.joinable()) enc->dispatcher.join();
    hvb_sync(enc->ctx);
    enc->arena.release();
}

struct Clock
{
    hvbenc *enc;
    int kind;
    std::chrono::steady_clock::time_point t0;
    Clock(hvbenc *e, int k) : enc(e), kind(k), t0(std::chrono::steady_clock::now()) {}
    ~Clock()
    {
        enc->waitNs[kind] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        enc->waitCount[kind] += 1;
    }
};

} // namespace

extern "C" int hvbenc_create(int device, int bytes_per_sample, int bit_depth, int width, int height, int pool_pictures, hvbenc **out)
{
    if (!out || pool_pictures < 2 || pool_pictures > 900) return HVB_ERR_INVALID;
    *out = nullptr;
    // every engine has its own stream; the driver maps streams onto this many hardware queues (default 8), and streams that
    // share a queue serialise.  Only effective before the process's first CUDA call, and never overrides the user's value.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int nEngines = 8;
    if (const char *v = getenv("HVB_ENGINES")) nEngines = std::max(1, std::min(64, atoi(v)));
    hvbenc *enc = new hvbenc;
    enc->bps = bytes_per_sample;
    for (int k = 0; k < 6; ++k)
    {
        enc->waitNs[k] = 0, enc->waitCount[k] = 0;
        for (int j = 0; j < 3; ++j) enc->phaseNs[k][j] = 0;
    }
    int rc = 0;
    const auto tStart = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart).count(); };
    for (int e = 0; e < nEngines && !rc; ++e)
    {
        Engine *engine = nullptr;
        rc = createEngine(device, bytes_per_sample, bit_depth, width, height, &engine);
        if (engine) enc->engines.push_back(engine);
    }
    const double tEngines = since();
    // the pictures: owned by engine 0, imported by the others in the same order, so an id means the same picture everywhere
    enc->slots.resize(pool_pictures);
    if (!rc) rc = hvb_picture_reserve(enc->engines[0]->ctx, width, height, 96, pool_pictures);
    for (int i = 0; i < pool_pictures && !rc; ++i)
    {
        rc = hvb_picture_create(enc->engines[0]->ctx, width, height, 96, &enc->slots[i].pic);
        for (size_t e = 1; e < enc->engines.size() && !rc; ++e)
        {
            int id = -1;
            rc = hvb_picture_import(enc->engines[e]->ctx, enc->engines[0]->ctx, enc->slots[i].pic, &id);
            if (!rc && id != enc->slots[i].pic) rc = HVB_ERR_INVALID;
        }
    }
    const double tPictures = since();
    enc->partition();
    if (enc->partitioned)
        for (size_t e = 0; e < enc->engines.size(); ++e)
            for (int k = 0; k < 6; ++k) enc->engines[e]->serves[k] = (int)e >= enc->first[k] && (int)e < enc->first[k + 1];
    for (size_t e = 0; e < enc->engines.size() && !rc; ++e) rc = equipEngine(enc->engines[e]);
    if (getenv("HVB_STATS") && atoi(getenv("HVB_STATS")))
        fprintf(stderr, "hvbenc set-up: %d contexts %.3f s, %d pictures %.3f s, page-locked buffers %.3f s\n", nEngines, tEngines, pool_pictures,
                tPictures - tEngines, since() - tPictures);
    if (rc)
    {
        fprintf(stderr, "hvbenc_create failed (%d): %s\n", rc, enc->engines.empty() ? "no engine" : hvb_last_error(enc->engines.back()->ctx));
        hvbenc_destroy(enc);
        return rc;
    }
    if (const char *v = getenv("HVB_POLLER")) enc->usePoller = atoi(v) != 0;
    if (const char *v = getenv("HVB_POLL_SLEEP_US")) enc->pollSleepUs = atoi(v);
    int nServices = 0; // measured (B200 call 20): a dispatcher thread per engine answers sooner than four threads sweeping eight engines each
    if (const char *v = getenv("HVB_SERVICE_THREADS")) nServices = std::max(0, std::min(16, atoi(v)));
    for (Engine *engine : enc->engines)
    {
        engine->session = enc;
        if (!engine->doneFlag) nServices = 0; // (HVB_DONE_FLAG=0: the flag is what the service threads watch)
    }
    if (nServices > 0)
    {
        nServices = std::min(nServices, (int)enc->engines.size());
        for (int k = 0; k < nServices; ++k) enc->services.emplace_back(serviceLoop, enc, k, nServices);
    }
    else
    {
        enc->poller = std::thread(pollLoop, enc);
        for (Engine *engine : enc->engines) engine->dispatcher = std::thread(dispatch, engine);
    }
    *out = enc;
    return HVB_OK;
}

extern "C" void hvbenc_destroy(hvbenc *enc)
{
    if (!enc) return;
    if (!enc->services.empty())
    {
        // let the batches in flight come back and the gathered requests go out, then stop sweeping
        for (int spin = 0; spin < 20000; ++spin)
        {
            bool busy = false;
            for (Engine *engine : enc->engines) busy |= engine->inflight.load() > 0;
            if (!busy) break;
            std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
        enc->serviceStop.store(true, std::memory_order_release);
        for (std::thread &t : enc->services) t.join();
    }
    for (Engine *engine : enc->engines) destroyEngine(engine);
    {
        std::lock_guard<std::mutex> g(enc->pollM);
        enc->pollStop = true;
    }
    enc->pollCv.notify_all();
    if (enc->poller.joinable()) enc->poller.join();
    // importing contexts first, the owner of the pictures last
    for (size_t e = enc->engines.size(); e-- > 0;)
    {
        hvb_destroy(enc->engines[e]->ctx);
        delete enc->engines[e];
    }
    delete enc;
}

extern "C" const char *hvbenc_last_error(hvbenc *enc)
{
    if (!enc) return "null session";
    for (Engine *engine : enc->engines)
        if (!engine->lastError.empty()) return engine->lastError.c_str();
    return enc->lastError.c_str();
}

extern "C" int hvbenc_picture(hvbenc *enc, const void *key, int fresh, int *pic)
{
    if (!enc || !key || !pic) return HVB_ERR_INVALID;
    std::lock_guard<std::mutex> g(enc->poolMutex);
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

extern "C" int hvbenc_upload_rects(hvbenc *session, int pic, const hvbenc_rect *rects, int n)
{
    if (!session || !rects || n <= 0) return HVB_ERR_INVALID;
    size_t total = 0;
    for (int i = 0; i < n; ++i)
    {
        if (!rects[i].host || rects[i].w <= 0 || rects[i].h <= 0 || rects[i].stride < rects[i].w) return HVB_ERR_INVALID;
        total += ((size_t)rects[i].w * session->bps * rects[i].h + 63) & ~size_t(63);
    }
    if (total > kUploadBytes / 2)
    {
        // more than the staging buffer affords in one piece: rectangle by rectangle, split by rows
        for (int i = 0; i < n; ++i)
        {
            const hvbenc_rect &r = rects[i];
            const size_t rowBytes = (size_t)r.w * session->bps;
            const int rows = std::max(1, (int)(kUploadBytes / 4 / rowBytes));
            for (int y = 0; y < r.h; y += rows)
            {
                hvbenc_rect part = r;
                part.host = static_cast<const char *>(r.host) + (size_t)y * r.stride * session->bps;
                part.y0 = r.y0 + y;
                part.h = std::min(rows, r.h - y);
                const int rc = hvbenc_upload_rects(session, pic, &part, 1);
                if (rc) return rc;
            }
        }
        return HVB_OK;
    }
    Clock clock(session, 0);
    Engine *enc = session->pick(0);
    Waiter local;
    Waiter &wt = armWaiter(local);
    wt.tSubmit = nowNs();
    enc->inflight.fetch_add(1, std::memory_order_relaxed);
    {
        std::unique_lock<std::mutex> lock(enc->m);
        enc->spaceCv.wait(lock, [&] { return enc->uploadUsed[enc->fill] + total <= kUploadBytes; });
        const int b = enc->fill;
        for (int i = 0; i < n; ++i)
        {
            const hvbenc_rect &r = rects[i];
            const size_t rowBytes = (size_t)r.w * enc->bps;
            char *dst = enc->uploadStage[b] + enc->uploadUsed[b];
            for (int y = 0; y < r.h; ++y) memcpy(dst + y * rowBytes, static_cast<const char *>(r.host) + (size_t)y * r.stride * enc->bps, rowBytes);
            enc->uploads[b].push_back(UploadTask{pic, r.cIdx, r.x0, r.y0, r.w, r.h, enc->uploadUsed[b]});
            enc->uploadUsed[b] += (rowBytes * r.h + 63) & ~size_t(63);
        }
        enc->uploadWaiters[b].push_back(&wt);
        ++enc->pending;
    }
    enc->workCv.notify_one();
    const int rcWait = waitFor(wt);
    notePhases(session, 0, wt);
    return rcWait;
}

extern "C" int hvbenc_upload_rect(hvbenc *session, int pic, int cIdx, const void *host, intptr_t stride, int x0, int y0, int w, int h)
{
    const hvbenc_rect r = {cIdx, host, stride, x0, y0, w, h};
    return hvbenc_upload_rects(session, pic, &r, 1);
}

extern "C" int hvbenc_me(hvbenc *session, const hvb_me_task *task, hvb_me_result *out)
{
    if (!session) return HVB_ERR_INVALID;
    Clock clock(session, 1);
    Engine *enc = session->pick(1);
    return submit(enc, 1, enc->me, task, 1, out, [](int, int) { return true; });
}

extern "C" int hvbenc_me_bi(hvbenc *session, const hvb_me_bi_task *task, hvb_me_bi_result *out)
{
    if (!session) return HVB_ERR_INVALID;
    Clock clock(session, 2);
    Engine *enc = session->pick(2);
    return submit(enc, 2, enc->bi, task, 1, out, [](int, int) { return true; });
}

extern "C" int hvbenc_pu_cost(hvbenc *session, const hvb_pu_cost_task *tasks, int n, int32_t *out)
{
    if (!session) return HVB_ERR_INVALID;
    Clock clock(session, 3);
    Engine *enc = session->pick(3);
    return submit(enc, 3, enc->pu, tasks, n, out, [](int, int) { return true; });
}

extern "C" int hvbenc_intra_sweep(hvbenc *session, const hvb_intra_sweep_task *task, const void *neighbours, int32_t *out)
{
    if (!session || !task || !neighbours || task->log2n < 2 || task->log2n > 5) return HVB_ERR_INVALID;
    Clock clock(session, 4);
    Engine *enc = session->pick(4);
    const size_t count = (size_t)(4 << task->log2n) + 1, room = (count + 15) & ~size_t(15);
    return submit(enc, 4, enc->intra, task, 1, out, [&](int b, int first) {
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

extern "C" int hvbenc_tu_chain(hvbenc *session, hvb_tu_task *tasks, int n, const hvb_rdoq_ctx *snapshot, const void *const *pred,
                               const intptr_t *pred_stride, void *const *rec, const intptr_t *rec_stride, int16_t *const *levels, hvb_tu_result *out)
{
    if (!session || !tasks || n <= 0 || !pred || !pred_stride || !rec || !rec_stride || !levels || !out) return HVB_ERR_INVALID;
    Clock clock(session, 5);
    Engine *enc = session->pick(5);
    return submit(enc, 5, enc->tu, tasks, n, out, [&](int b, int first) {
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

extern "C" void hvbenc_set_thread_hooks(const hvbenc_thread_hooks *hooks)
{
    tlHooks = hooks ? *hooks : hvbenc_thread_hooks{nullptr, nullptr, nullptr};
}

extern "C" int hvbenc_stats(hvbenc *session, char *buf, size_t bytes)
{
    if (!session || !buf || !bytes) return HVB_ERR_INVALID;
    long long dispatches = 0, launches = 0, tasks[5] = {0, 0, 0, 0, 0}, batches[5] = {0, 0, 0, 0, 0}, rects = 0, rectBytes = 0;
    double busy = 0, kindMs[6] = {0, 0, 0, 0, 0, 0}, alg[6] = {0, 0, 0, 0, 0, 0}, toDevice = 0, fromDevice = 0;
    for (Engine *enc : session->engines)
    {
        std::lock_guard<std::mutex> g(enc->m);
        for (int k = 0; k < 6; ++k) kindMs[k] += enc->kindMs[k], alg[k] += enc->algBytes[k];
        toDevice += enc->toDevice;
        fromDevice += enc->fromDevice;
        dispatches += enc->dispatches;
        busy += enc->deviceSeconds;
        launches += hvb_launch_count(enc->ctx);
        const long long t[5] = {enc->me.totalTasks, enc->bi.totalTasks, enc->pu.totalTasks, enc->intra.totalTasks, enc->tu.totalTasks};
        const long long bt[5] = {enc->me.totalBatches, enc->bi.totalBatches, enc->pu.totalBatches, enc->intra.totalBatches, enc->tu.totalBatches};
        for (int k = 0; k < 5; ++k) tasks[k] += t[k], batches[k] += bt[k];
        rects += enc->totalUploads;
        rectBytes += enc->totalUploadBytes;
    }
    const char *names[5] = {"me", "me_bi", "pu_cost", "intra_sweep", "tu_chain"};
    int at = snprintf(buf, bytes, "{\"engines\": %d, \"dispatches\": %lld, \"engine_busy_s\": %.3f, \"kernel_launches\": %lld, ", (int)session->engines.size(),
                      dispatches, busy, launches);
    for (int k = 0; k < 5 && at < (int)bytes; ++k)
    {
        const long long c = session->waitCount[k + 1];
        at += snprintf(buf + at, bytes - at,
                       "\"%s\": {\"tasks\": %lld, \"batches\": %lld, \"requests\": %lld, \"mean_wait_us\": %.1f, \"device_ms\": %.3f, "
                       "\"algorithmic_bytes\": %.0f, \"mean_queue_us\": %.1f, \"mean_batch_us\": %.1f, \"mean_wake_us\": %.1f}, ",
                       names[k], tasks[k], batches[k], c, c ? session->waitNs[k + 1] / 1000.0 / c : 0.0, kindMs[k + 1], alg[k + 1],
                       c ? session->phaseNs[k + 1][0] / 1000.0 / c : 0.0, c ? session->phaseNs[k + 1][1] / 1000.0 / c : 0.0,
                       c ? session->phaseNs[k + 1][2] / 1000.0 / c : 0.0);
    }
    const long long uc = session->waitCount[0];
    if (at < (int)bytes)
        snprintf(buf + at, bytes - at,
                 "\"uploads\": {\"rects\": %lld, \"bytes\": %lld, \"mean_wait_us\": %.1f, \"device_ms\": %.3f}, "
                 "\"h2d_bytes\": %.0f, \"d2h_bytes\": %.0f}",
                 rects, rectBytes, uc ? session->waitNs[0] / 1000.0 / uc : 0.0, kindMs[0], toDevice, fromDevice);
    return HVB_OK;
}
