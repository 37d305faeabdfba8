// hvb_context.cu -- context, device-resident pictures, pools and host<->device staging.
//
// Replaces the reference's havoc_code life cycle (havoc/havoc.h:138-149, havoc.cpp:144-155) and gives
// the batched kernels what the reference gets from plain CPU pointers: source pictures
// (turing/encode.cpp:363-451 copies planes into PictureWrap) and padded reconstructed pictures
// (turing/StatePictures.h:154-156, Padding.h) living in HBM.
#include "hvb_internal.cuh"
#include <mutex>
#include <utility>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

static_assert(sizeof(hvb_block) == 8, "abi");
static_assert(sizeof(hvb_metric_task) == 24, "abi");
static_assert(sizeof(hvb_sad4_task) == 32, "abi");
static_assert(sizeof(hvb_pred_task) == 32, "abi");
static_assert(sizeof(hvb_subtract_bi_task) == 32, "abi");
static_assert(sizeof(hvb_interp_satd_task) == 20, "abi");
static_assert(sizeof(hvb_intra_task) == 16, "abi");
static_assert(sizeof(hvb_intra_sweep_task) == 24, "abi");
static_assert(sizeof(hvb_transform_task) == 16, "abi");
static_assert(sizeof(hvb_quant_task) == 24, "abi");
static_assert(sizeof(hvb_ita_task) == 24, "abi");
static_assert(sizeof(hvb_tu_task) == 60, "abi");
static_assert(sizeof(hvb_tu_result) == 32, "abi");
static_assert(sizeof(hvb_rdoq_ctx) == 136, "abi");
static_assert(sizeof(hvb_rdoq_task) == 28, "abi");
static_assert(sizeof(hvb_me_task) == 64, "abi");
static_assert(sizeof(hvb_me_result) == 56, "abi");
static_assert(sizeof(hvb_me_bi_task) == 64, "abi");
static_assert(sizeof(hvb_me_bi_result) == 32, "abi");
static_assert(sizeof(hvb_pu_cost_task) == 24, "abi");

int hvbFail(hvb_context *ctx, int status, const char *what)
{
    if (ctx) ctx->lastError = what;
    return status;
}

int hvbCuda(hvb_context *ctx, cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return HVB_OK;
    if (ctx)
    {
        ctx->lastError = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    }
    return e == cudaErrorMemoryAllocation ? HVB_ERR_NOMEM : HVB_ERR_CUDA;
}

extern "C" int hvb_device_ok(int device)
{
    // asked once per device: cudaGetDeviceProperties takes tens of milliseconds, and a session creates dozens of contexts
    static std::mutex m;
    static int known[64]; // 0: not asked, 1: usable, 2: not
    if (device < 0) return 0;
    std::lock_guard<std::mutex> g(m);
    if (device < 64 && known[device]) return known[device] == 1;
    int count = 0, major = 0;
    int ok = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess && device < count &&
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) == cudaSuccess)
        ok = major == 10; // kernels are compiled for sm_100a only
    if (device < 64) known[device] = ok ? 1 : 2;
    return ok;
}

extern "C" int hvb_create(int device, int bytes_per_sample, int bit_depth, hvb_context **out)
{
    if (!out) return HVB_ERR_INVALID;
    *out = nullptr;
    if (bytes_per_sample != 1 && bytes_per_sample != 2) return HVB_ERR_INVALID;
    if (bit_depth < 8 || bit_depth > (bytes_per_sample == 1 ? 8 : 10)) return HVB_ERR_INVALID;
    // No CPU fallback: without a Blackwell device the product path fails loudly.
    if (!hvb_device_ok(device)) return HVB_ERR_NO_DEVICE;

    hvb_context *ctx = new (std::nothrow) hvb_context;
    if (!ctx) return HVB_ERR_NOMEM;
    ctx->device = device;
    ctx->bps = bytes_per_sample;
    ctx->bitDepth = bit_depth;

    cudaError_t e = cudaSetDevice(device);
    if (const char *v = getenv("HVB_BLOCKING_SYNC"))
        if (atoi(v))
        {
            // a thread that waits for a stream sleeps until the device's interrupt instead of spinning in the driver: what a
            // session with dozens of dispatcher threads on a few cores wants (hvb_encoder.cpp, HVB_POLLER=0)
            cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
            cudaGetLastError();
        }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dPlanes, sizeof(HvbPlane) * HVB_MAX_PICTURES * 3);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->dPlanes, 0, sizeof(HvbPlane) * HVB_MAX_PICTURES * 3, ctx->ownStream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->ownStream);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->workCursors, 64 * sizeof(int));
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess)
    {
        hvb_destroy(ctx);
        return HVB_ERR_CUDA;
    }
    ctx->smCount = sms;
    ctx->useTma = hvbEnvTma();
    if (const char *v = getenv("HVB_TU_FUSED_MAX")) ctx->tuFusedMax = atoi(v) < 0 ? 0 : atoi(v);
    ctx->stream = ctx->ownStream;
    *out = ctx;
    return HVB_OK;
}

extern "C" void hvb_destroy(hvb_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->pictures)
    {
        for (int c = 0; c < 3; ++c)
            if (p.alloc[c]) cudaFree(p.alloc[c]);
        if (p.lfInfo) cudaFree(p.lfInfo);
        if (p.saoInfo) cudaFree(p.saoInfo);
    }
    if (ctx->slab) cudaFree(ctx->slab);
    if (ctx->dPlanes) cudaFree(ctx->dPlanes);
    if (ctx->dTensorMaps) cudaFree(ctx->dTensorMaps);
    if (ctx->dLoopInfo) cudaFree(ctx->dLoopInfo);
    if (ctx->workCursors) cudaFree(ctx->workCursors);
    for (auto &slot : ctx->slots)
    {
        if (slot.dev) cudaFree(slot.dev);
        if (slot.done) cudaEventDestroy(slot.done);
    }
    for (cudaEvent_t &m : ctx->marks)
        if (m) cudaEventDestroy(m);
    if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
    if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
    if (ctx->evIn) cudaEventDestroy(ctx->evIn);
    if (ctx->evCompute) cudaEventDestroy(ctx->evCompute);
    if (ctx->hostStage) cudaFreeHost(ctx->hostStage);
    if (ctx->devStage) cudaFree(ctx->devStage);
    if (ctx->samplePool) cudaFree(ctx->samplePool);
    if (ctx->coeffPool) cudaFree(ctx->coeffPool);
    if (ctx->rdoqCtx) cudaFree(ctx->rdoqCtx);
    if (ctx->rdoqBits) cudaFree(ctx->rdoqBits);
    if (ctx->rdoqLast) cudaFree(ctx->rdoqLast);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    delete ctx;
}

extern "C" const char *hvb_last_error(hvb_context *ctx) { return ctx ? ctx->lastError.c_str() : "null context"; }

extern "C" int hvb_set_stream(hvb_context *ctx, void *cuda_stream)
{
    if (!ctx) return HVB_ERR_INVALID;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->ownStream;
    return HVB_OK;
}

extern "C" int hvb_sync(hvb_context *ctx)
{
    if (!ctx) return HVB_ERR_INVALID;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && ctx->copyIn) e = cudaStreamSynchronize(ctx->copyIn);
    if (e == cudaSuccess && ctx->copyOut) e = cudaStreamSynchronize(ctx->copyOut);
    for (auto &slot : ctx->slots) slot.busy = false;
    return hvbCuda(ctx, e, "hvb_sync");
}

extern "C" int hvb_mark(hvb_context *ctx, int slot)
{
    HVB_CHECK_ARGS(ctx, slot >= 0 && slot < 16);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaSuccess;
    if (!ctx->marks[slot]) e = cudaEventCreate(&ctx->marks[slot]);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->marks[slot], ctx->stream);
    return hvbCuda(ctx, e, "hvb_mark");
}

extern "C" int hvb_elapsed_ms(hvb_context *ctx, int from, int to, float *ms)
{
    HVB_CHECK_ARGS(ctx, ms && from >= 0 && from < 16 && to >= 0 && to < 16 && ctx->marks[from] && ctx->marks[to]);
    return hvbCuda(ctx, cudaEventElapsedTime(ms, ctx->marks[from], ctx->marks[to]), "hvb_elapsed_ms");
}

extern "C" int hvb_set_tma(hvb_context *ctx, int on)
{
    if (!ctx) return HVB_ERR_INVALID;
    ctx->useTma = on != 0;
    return HVB_OK;
}

namespace {
__global__ void signalKernel(volatile int32_t *flag, int32_t value)
{
    __threadfence_system(); // everything the stream did before is visible to the host before the flag is
    *flag = value;
}
} // namespace

extern "C" int hvb_signal(hvb_context *ctx, int32_t *flag, int32_t value)
{
    HVB_CHECK_ARGS(ctx, flag);
    cudaSetDevice(ctx->device);
    signalKernel<<<1, 1, 0, ctx->stream>>>(flag, value);
    HVB_LAUNCH_CHECK(ctx, "signalKernel");
    return HVB_OK;
}

extern "C" int hvb_poll(hvb_context *ctx)
{
    if (!ctx) return HVB_ERR_INVALID;
    const cudaStream_t streams[3] = {ctx->stream, ctx->copyIn, ctx->copyOut};
    for (cudaStream_t s : streams)
    {
        if (!s) continue;
        const cudaError_t e = cudaStreamQuery(s);
        if (e == cudaErrorNotReady) return 0;
        if (e != cudaSuccess) return HVB_ERR_CUDA; // (lastError is the owning thread's to write)
    }
    return 1;
}

extern "C" int hvb_set_tu_fused_max(hvb_context *ctx, int blocks)
{
    HVB_CHECK_ARGS(ctx, blocks >= 0);
    ctx->tuFusedMax = blocks;
    return HVB_OK;
}

extern "C" int hvb_set_pipelined(hvb_context *ctx, int on)
{
    if (!ctx) return HVB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    int rc = hvb_sync(ctx);
    if (rc) return rc;
    if (on && !ctx->copyIn)
    {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evIn, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->evCompute, cudaEventDisableTiming);
        for (auto &slot : ctx->slots)
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "hvb_set_pipelined");
    }
    ctx->pipelined = on != 0;
    return HVB_OK;
}

int hvbUpload(hvb_context *ctx, void *dev, size_t devPitch, const void *host, size_t hostPitch, size_t widthBytes, size_t rows,
              const char *what)
{
    cudaSetDevice(ctx->device);
    if (ctx->pipelined && hvbIsPinned(host))
    {
        // Not ordered against batches issued EARLIER: like any asynchronous copy, the caller must not overwrite what work
        // in flight still reads (an encoder uploads the next frame into a free picture while the previous one is searched).
        cudaError_t e = cudaMemcpy2DAsync(dev, devPitch, host, hostPitch, widthBytes, rows, cudaMemcpyHostToDevice, ctx->copyIn);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->evIn, ctx->copyIn);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->evIn, 0);
        return hvbCuda(ctx, e, what);
    }
    cudaError_t e = cudaMemcpy2DAsync(dev, devPitch, host, hostPitch, widthBytes, rows, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // the host buffer may be pageable
    return hvbCuda(ctx, e, what);
}

bool hvbEnvTma()
{
    static const bool value = [] {
        const char *v = getenv("HVB_TMA");
        return v && *v && *v != '0';
    }();
    return value;
}

extern "C" int64_t hvb_launch_count(hvb_context *ctx) { return ctx ? ctx->launches : 0; }

// Page-locked ranges handed out by hvb_host_alloc: hvbIsPinned answers for them without asking the driver (every driver
// call takes the context's lock, and a session with dozens of engines lives on that lock's throughput).
namespace {
std::mutex gPinnedMutex;
std::vector<std::pair<uintptr_t, size_t>> gPinned;
} // namespace

extern "C" int hvb_host_alloc(hvb_context *ctx, size_t bytes, void **out)
{
    HVB_CHECK_ARGS(ctx, out && bytes > 0);
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int rc = hvbCuda(ctx, cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped), "hvb_host_alloc");
    if (!rc)
    {
        std::lock_guard<std::mutex> g(gPinnedMutex);
        gPinned.emplace_back(reinterpret_cast<uintptr_t>(*out), bytes);
    }
    return rc;
}

extern "C" int hvb_host_free(hvb_context *ctx, void *ptr)
{
    HVB_CHECK_ARGS(ctx, ptr);
    cudaSetDevice(ctx->device);
    {
        std::lock_guard<std::mutex> g(gPinnedMutex);
        for (size_t i = 0; i < gPinned.size(); ++i)
            if (gPinned[i].first == reinterpret_cast<uintptr_t>(ptr))
            {
                gPinned.erase(gPinned.begin() + i);
                break;
            }
    }
    return hvbCuda(ctx, cudaFreeHost(ptr), "hvb_host_free");
}

// ---------------------------------------------------------------------------------------------
// Pictures
// ---------------------------------------------------------------------------------------------

extern "C" int hvb_picture_create(hvb_context *ctx, int width, int height, int pad, int *pic)
{
    HVB_CHECK_ARGS(ctx, pic && width > 0 && height > 0 && pad >= 0 && !(width & 1) && !(height & 1) && !(pad & 1));
    int id = -1;
    for (int i = 0; i < HVB_MAX_PICTURES; ++i)
        if (!ctx->pictures[i].live)
        {
            id = i;
            break;
        }
    if (id < 0) return hvbFail(ctx, HVB_ERR_NOMEM, "picture table full");
    cudaSetDevice(ctx->device);
    HvbPicture &p = ctx->pictures[id];
    p.width = width;
    p.height = height;
    p.pad = pad;
    for (int c = 0; c < 3; ++c)
    {
        const int w = c ? width / 2 : width, h = c ? height / 2 : height, pd = c ? pad / 2 : pad;
        // row pitch: multiple of 256 bytes; left padding rounded up so that sample (0,0) of each row is
        // 256-byte aligned (source rows are then always 128-bit loadable).
        const size_t padLeft = ((size_t)pd * ctx->bps + 255) / 256 * 256 / ctx->bps;
        const size_t strideSamples = (padLeft + w + pd + 16 + (256 / ctx->bps - 1)) / (256 / ctx->bps) * (256 / ctx->bps);
        const size_t rows = (size_t)h + 2 * pd + 2; // +2 slack rows: vector loads may run a few bytes past a block
        const size_t bytes = strideSamples * rows * ctx->bps;
        // from the context's reserve when there is one (hvb_picture_reserve: one allocation for a whole pool of pictures; the
        // plane is then not the picture's to free), else its own allocation
        void *mem = nullptr;
        const size_t need = (bytes + 255) & ~size_t(255);
        if (ctx->slab && ctx->slabUsed + need <= ctx->slabBytes)
        {
            mem = static_cast<char *>(ctx->slab) + ctx->slabUsed;
            ctx->slabUsed += need;
            p.alloc[c] = nullptr;
        }
        cudaError_t e = mem ? cudaSuccess : cudaMalloc(&p.alloc[c], bytes);
        if (!mem) mem = p.alloc[c];
        if (e != cudaSuccess)
        {
            for (int k = 0; k < c; ++k)
            {
                cudaFree(p.alloc[k]);
                p.alloc[k] = nullptr;
            }
            return hvbCuda(ctx, e, "hvb_picture_create");
        }
        if (p.alloc[c]) cudaMemsetAsync(mem, 0, bytes, ctx->stream); // (the reserve was cleared when it was made)
        p.allocBytes[c] = bytes;
        p.plane[c].base = static_cast<char *>(mem) + ((size_t)pd * strideSamples + padLeft) * ctx->bps;
        p.plane[c].stride = (int32_t)strideSamples;
        p.plane[c].width = w;
        p.plane[c].height = h;
        p.plane[c].pad = pd;
        p.plane[c].reserved = (int32_t)padLeft; // sample (-pd) of a row starts padLeft - pd samples into it; (padLeft + x) indexes the row
        p.tmaBase[c] = mem;
        p.tmaRows[c] = (int)rows;
    }
    p.live = true;
    ctx->planesDirty = true;
    ctx->tensorMapsDirty = true;
    *pic = id;
    return HVB_OK;
}

extern "C" int hvb_picture_reserve(hvb_context *ctx, int width, int height, int pad, int count)
{
    HVB_CHECK_ARGS(ctx, width > 0 && height > 0 && pad >= 0 && !(width & 1) && !(height & 1) && !(pad & 1) && count > 0 && !ctx->slab);
    cudaSetDevice(ctx->device);
    size_t perPicture = 0;
    for (int c = 0; c < 3; ++c)
    {
        // (the geometry of hvb_picture_create)
        const int w = c ? width / 2 : width, h = c ? height / 2 : height, pd = c ? pad / 2 : pad;
        const size_t padLeft = ((size_t)pd * ctx->bps + 255) / 256 * 256 / ctx->bps;
        const size_t strideSamples = (padLeft + w + pd + 16 + (256 / ctx->bps - 1)) / (256 / ctx->bps) * (256 / ctx->bps);
        const size_t rows = (size_t)h + 2 * pd + 2;
        perPicture += (strideSamples * rows * ctx->bps + 255) & ~size_t(255);
    }
    const size_t bytes = perPicture * (size_t)count;
    cudaError_t e = cudaMalloc(&ctx->slab, bytes);
    if (e != cudaSuccess)
    {
        ctx->slab = nullptr;
        return hvbCuda(ctx, e, "hvb_picture_reserve");
    }
    ctx->slabBytes = bytes;
    ctx->slabUsed = 0;
    return hvbCuda(ctx, cudaMemsetAsync(ctx->slab, 0, bytes, ctx->stream), "hvb_picture_reserve");
}

extern "C" int hvb_picture_wrap(hvb_context *ctx, void *host, intptr_t stride, int width, int height, int *pic)
{
    HVB_CHECK_ARGS(ctx, pic && host && width > 0 && height > 0 && stride >= width);
    int id = -1;
    for (int i = 0; i < HVB_MAX_PICTURES; ++i)
        if (!ctx->pictures[i].live)
        {
            id = i;
            break;
        }
    if (id < 0) return hvbFail(ctx, HVB_ERR_NOMEM, "picture table full");
    cudaSetDevice(ctx->device);
    void *dev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dev, host, 0);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "hvb_picture_wrap: not memory from hvb_host_alloc");
    HvbPicture &p = ctx->pictures[id];
    p = HvbPicture{};
    p.width = width;
    p.height = height;
    p.pad = 0;
    for (int c = 0; c < 3; ++c) p.plane[c] = HvbPlane{nullptr, 0, 0, 0, 0, 0};
    p.plane[0] = HvbPlane{dev, (int32_t)stride, width, height, 0, 0};
    p.live = true;
    ctx->planesDirty = true;
    *pic = id;
    return HVB_OK;
}

extern "C" int hvb_picture_import(hvb_context *ctx, hvb_context *owner, int owner_pic, int *pic)
{
    HVB_CHECK_ARGS(ctx, pic && owner && owner != ctx && owner->device == ctx->device && owner->bps == ctx->bps && owner_pic >= 0 &&
                            owner_pic < HVB_MAX_PICTURES && owner->pictures[owner_pic].live);
    int id = -1;
    for (int i = 0; i < HVB_MAX_PICTURES; ++i)
        if (!ctx->pictures[i].live)
        {
            id = i;
            break;
        }
    if (id < 0) return hvbFail(ctx, HVB_ERR_NOMEM, "picture table full");
    HvbPicture &p = ctx->pictures[id];
    p = owner->pictures[owner_pic];
    for (int c = 0; c < 3; ++c) p.alloc[c] = nullptr; // borrowed: hvb_picture_destroy / hvb_destroy free nothing
    p.lfInfo = p.saoInfo = nullptr;
    p.lfBytes = p.saoBytes = 0;
    ctx->planesDirty = true;
    ctx->tensorMapsDirty = true;
    *pic = id;
    return HVB_OK;
}

extern "C" int hvb_picture_destroy(hvb_context *ctx, int pic)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    HvbPicture &p = ctx->pictures[pic];
    for (int c = 0; c < 3; ++c)
    {
        if (p.alloc[c]) cudaFree(p.alloc[c]); // a wrapped picture owns nothing
        p.alloc[c] = nullptr;
    }
    if (p.lfInfo || p.saoInfo)
    {
        // the device table entry goes with it: a later hvb_deblock_batch on a recycled id must find no stale pointers
        if (p.lfInfo) cudaFree(p.lfInfo);
        if (p.saoInfo) cudaFree(p.saoInfo);
        p.lfInfo = p.saoInfo = nullptr;
        p.lfBytes = p.saoBytes = 0;
        ctx->loopInfoHost[pic] = HvbLoopInfo{};
        if (ctx->dLoopInfo)
        {
            cudaMemsetAsync(ctx->dLoopInfo + pic, 0, sizeof(HvbLoopInfo), ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
    }
    p.live = false;
    ctx->planesDirty = true;
    return HVB_OK;
}

int hvbSyncPlanes(hvb_context *ctx)
{
    if (!ctx->planesDirty) return HVB_OK;
    std::vector<HvbPlane> host(HVB_MAX_PICTURES * 3);
    memset(host.data(), 0, host.size() * sizeof(HvbPlane));
    for (int i = 0; i < HVB_MAX_PICTURES; ++i)
        if (ctx->pictures[i].live)
            for (int c = 0; c < 3; ++c) host[i * 3 + c] = ctx->pictures[i].plane[c];
    // synchronous small copy: the table changes only when pictures are created or destroyed
    // (on the context's stream -- ordered against its kernels -- and waited for: `host` is pageable and local)
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ctx->dPlanes, host.data(), host.size() * sizeof(HvbPlane), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "hvbSyncPlanes");
    ctx->planesDirty = false;
    return HVB_OK;
}

static int planeCopy(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride, int y0, int rows, bool upload)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live && cIdx >= 0 && cIdx < 3 && host);
    const HvbPlane &pl = ctx->pictures[pic].plane[cIdx];
    HVB_CHECK_ARGS(ctx, y0 >= 0 && rows >= 0 && y0 + rows <= pl.height && stride >= pl.width);
    if (!rows) return HVB_OK;
    cudaSetDevice(ctx->device);
    char *dev = static_cast<char *>(pl.base) + (size_t)y0 * pl.stride * ctx->bps;
    char *h = static_cast<char *>(host) + (size_t)y0 * stride * ctx->bps;
    if (upload)
        return hvbUpload(ctx, dev, (size_t)pl.stride * ctx->bps, h, (size_t)stride * ctx->bps, (size_t)pl.width * ctx->bps, rows,
                         "hvb_picture_upload");
    cudaError_t e = cudaMemcpy2DAsync(h, (size_t)stride * ctx->bps, dev, (size_t)pl.stride * ctx->bps, (size_t)pl.width * ctx->bps, rows,
                                      cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // host buffer may be pageable
    return hvbCuda(ctx, e, "hvb_picture_download");
}

extern "C" int hvb_picture_upload(hvb_context *ctx, int pic, int cIdx, const void *host, intptr_t stride, int y0, int rows)
{
    return planeCopy(ctx, pic, cIdx, const_cast<void *>(host), stride, y0, rows, true);
}

extern "C" int hvb_picture_download(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride, int y0, int rows)
{
    return planeCopy(ctx, pic, cIdx, host, stride, y0, rows, false);
}

static int rectCopy(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride, int x0, int y0, int w, int h, bool upload)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live && cIdx >= 0 && cIdx < 3 && host);
    const HvbPlane &pl = ctx->pictures[pic].plane[cIdx];
    HVB_CHECK_ARGS(ctx, w >= 0 && h >= 0 && stride >= w && x0 >= -pl.pad && y0 >= -pl.pad && x0 + w <= pl.width + pl.pad &&
                            y0 + h <= pl.height + pl.pad);
    if (!w || !h) return HVB_OK;
    cudaSetDevice(ctx->device);
    char *dev = static_cast<char *>(pl.base) + ((intptr_t)y0 * pl.stride + x0) * ctx->bps;
    cudaError_t e;
    if (upload)
        return hvbUpload(ctx, dev, (size_t)pl.stride * ctx->bps, host, (size_t)stride * ctx->bps, (size_t)w * ctx->bps, h,
                         "hvb_picture_upload_rect");
    else
        e = cudaMemcpy2DAsync(host, (size_t)stride * ctx->bps, dev, (size_t)pl.stride * ctx->bps, (size_t)w * ctx->bps, h,
                              cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    return hvbCuda(ctx, e, upload ? "hvb_picture_upload_rect" : "hvb_picture_download_rect");
}

extern "C" int hvb_picture_upload_rect(hvb_context *ctx, int pic, int cIdx, const void *host, intptr_t stride, int x0, int y0, int w, int h)
{
    return rectCopy(ctx, pic, cIdx, const_cast<void *>(host), stride, x0, y0, w, h, true);
}

extern "C" int hvb_picture_download_rect(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride, int x0, int y0, int w, int h)
{
    return rectCopy(ctx, pic, cIdx, host, stride, x0, y0, w, h, false);
}

extern "C" int hvb_picture_plane(hvb_context *ctx, int pic, int cIdx, void **dev_ptr, intptr_t *stride)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live && cIdx >= 0 && cIdx < 3);
    if (dev_ptr) *dev_ptr = ctx->pictures[pic].plane[cIdx].base;
    if (stride) *stride = ctx->pictures[pic].plane[cIdx].stride;
    return HVB_OK;
}

// Edge replication (turing/Padding.h): every padding sample takes the value of the nearest picture sample.
template <typename Sample>
__global__ void padKernel(HvbPlane pl)
{
    const int W = pl.width + 2 * pl.pad, H = pl.height + 2 * pl.pad;
    Sample *base = reinterpret_cast<Sample *>(pl.base);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < W * H; idx += gridDim.x * blockDim.x)
    {
        const int x = idx % W - pl.pad, y = idx / W - pl.pad;
        if (x >= 0 && x < pl.width && y >= 0 && y < pl.height) continue;
        const int sx = min(max(x, 0), pl.width - 1), sy = min(max(y, 0), pl.height - 1);
        base[(intptr_t)y * pl.stride + x] = base[(intptr_t)sy * pl.stride + sx];
    }
}

extern "C" int hvb_picture_pad(hvb_context *ctx, int pic)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live);
    cudaSetDevice(ctx->device);
    for (int c = 0; c < 3; ++c)
    {
        const HvbPlane &pl = ctx->pictures[pic].plane[c];
        if (!pl.pad) continue;
        const int blocks = ctx->smCount * 4;
        if (ctx->bps == 1)
            padKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(pl);
        else
            padKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(pl);
        HVB_LAUNCH_CHECK(ctx, "padKernel");
    }
    return HVB_OK;
}

// ---------------------------------------------------------------------------------------------
// Pools and staging
// ---------------------------------------------------------------------------------------------

// Pool growth.  Everything is issued on the context's stream (the stream is cudaStreamNonBlocking or a caller's
// stream, so work on the legacy default stream would not be ordered against the kernels that follow) and the
// host waits for it before the old allocation is released.  preserve = false (kernel workspace): the old
// contents are neither copied nor is the new allocation cleared.
template <typename T>
static int growDevice(hvb_context *ctx, T **ptr, size_t *have, size_t want, size_t elemBytes, const char *what, bool preserve = true)
{
    if (*have >= want) return HVB_OK;
    cudaSetDevice(ctx->device);
    size_t cap = *have ? *have : 1024;
    while (cap < want) cap *= 2;
    void *fresh = nullptr;
    cudaError_t e = cudaStreamSynchronize(ctx->stream); // nothing in flight may still use the old allocation
    if (e == cudaSuccess && ctx->copyIn) e = cudaStreamSynchronize(ctx->copyIn);
    if (e == cudaSuccess && ctx->copyOut) e = cudaStreamSynchronize(ctx->copyOut);
    if (e != cudaSuccess) return hvbCuda(ctx, e, what);
    if (!preserve && *ptr)
    {
        cudaFree(*ptr);
        *ptr = nullptr;
        *have = 0;
    }
    e = cudaMalloc(&fresh, cap * elemBytes);
    if (e != cudaSuccess) return hvbCuda(ctx, e, what);
    if (preserve)
    {
        if (*ptr) e = cudaMemcpyAsync(fresh, *ptr, *have * elemBytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemsetAsync(static_cast<char *>(fresh) + *have * elemBytes, 0, (cap - *have) * elemBytes, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (*ptr) cudaFree(*ptr);
        if (e != cudaSuccess)
        {
            cudaFree(fresh);
            *ptr = nullptr;
            *have = 0;
            return hvbCuda(ctx, e, what);
        }
    }
    *ptr = static_cast<T *>(fresh);
    *have = cap;
    return HVB_OK;
}

int hvbEnsureScratch(hvb_context *ctx, size_t bytes)
{
    return growDevice(ctx, reinterpret_cast<char **>(&ctx->scratch), &ctx->scratchBytes, bytes, 1, "scratch", false);
}

int hvbEnsureCoeffPool(hvb_context *ctx, size_t count)
{
    return growDevice(ctx, &ctx->coeffPool, &ctx->coeffPoolCount, count, sizeof(int16_t), "coeff pool");
}

int hvbEnsureSamplePool(hvb_context *ctx, size_t count)
{
    return growDevice(ctx, reinterpret_cast<char **>(&ctx->samplePool), &ctx->samplePoolCount, count, (size_t)ctx->bps,
                      "sample pool");
}

extern "C" int hvb_pool_upload(hvb_context *ctx, const void *samples, size_t count, size_t offset)
{
    HVB_CHECK_ARGS(ctx, samples || !count);
    int rc = hvbEnsureSamplePool(ctx, offset + count + 64);
    if (rc) return rc;
    if (!count) return HVB_OK;
    return hvbUpload(ctx, static_cast<char *>(ctx->samplePool) + offset * ctx->bps, count * ctx->bps, samples, count * ctx->bps,
                     count * ctx->bps, 1, "hvb_pool_upload");
}

extern "C" int hvb_coeff_upload(hvb_context *ctx, const int16_t *data, size_t count, size_t offset)
{
    HVB_CHECK_ARGS(ctx, data || !count);
    int rc = hvbEnsureCoeffPool(ctx, offset + count);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(ctx->coeffPool + offset, data, count * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    return hvbCuda(ctx, e, "hvb_coeff_upload");
}

extern "C" int hvb_coeff_download(hvb_context *ctx, int16_t *data, size_t count, size_t offset)
{
    HVB_CHECK_ARGS(ctx, (data || !count) && offset + count <= ctx->coeffPoolCount);
    cudaError_t e = cudaMemcpyAsync(data, ctx->coeffPool + offset, count * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && !(ctx->pipelined && hvbIsPinned(data))) e = cudaStreamSynchronize(ctx->stream);
    return hvbCuda(ctx, e, "hvb_coeff_download");
}

extern "C" int hvb_coeff_pool_wrap(hvb_context *ctx, int16_t *levels, size_t count)
{
    HVB_CHECK_ARGS(ctx, (levels && count > 0 && hvbIsPinned(levels)) || (!levels && !count));
    ctx->coeffWrap = levels;
    ctx->coeffWrapCount = count;
    return HVB_OK;
}

extern "C" int hvb_rdoq_contexts_wrap(hvb_context *ctx, const hvb_rdoq_ctx *snapshots, int count)
{
    HVB_CHECK_ARGS(ctx, (snapshots && count > 0 && hvbIsPinned(snapshots)) || (!snapshots && !count));
    ctx->rdoqWrap = snapshots;
    ctx->rdoqWrapCount = count;
    return HVB_OK;
}

extern "C" int hvb_rdoq_contexts_upload(hvb_context *ctx, const hvb_rdoq_ctx *snapshots, int count, int first)
{
    HVB_CHECK_ARGS(ctx, snapshots && count > 0 && first >= 0);
    size_t have = (size_t)ctx->rdoqCtxCount;
    int rc = growDevice(ctx, &ctx->rdoqCtx, &have, (size_t)first + count, sizeof(hvb_rdoq_ctx), "rdoq contexts");
    if (rc) return rc;
    ctx->rdoqCtxCount = (int)have;
    rc = growDevice(ctx, &ctx->rdoqBits, &ctx->rdoqBitsCount, have * sizeof(hvb_rdoq_ctx), sizeof(int2), "rdoq bit costs");
    if (rc) return rc;
    rc = growDevice(ctx, &ctx->rdoqLast, &ctx->rdoqLastCount, have * 160, sizeof(int), "rdoq last-position rates");
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(ctx->rdoqCtx + first, snapshots, sizeof(hvb_rdoq_ctx) * count, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "hvb_rdoq_contexts_upload");
    rc = hvbLaunchRdoqBits(ctx, first, count); // {bits(0), bits(1)} per state byte: what the RDOQ walk actually reads
    if (rc) return rc;
    if (ctx->pipelined && hvbIsPinned(snapshots)) return HVB_OK; // the source stays the library's until hvb_sync
    return hvbCuda(ctx, cudaStreamSynchronize(ctx->stream), "hvb_rdoq_contexts_upload");
}

bool hvbIsPinned(const void *p)
{
    if (p)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        std::lock_guard<std::mutex> g(gPinnedMutex);
        for (const auto &r : gPinned)
            if (a >= r.first && a - r.first < r.second) return true;
    }
    cudaPointerAttributes attr;
    if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess)
    {
        cudaGetLastError(); // unregistered host memory reports an error on older runtimes: clear it
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int hvbStageIn(hvb_context *ctx, const void *tasks, size_t inBytes, void *out, size_t outBytes, hvb_mem mem, HvbStaged *st)
{
    int rc = hvbSyncPlanes(ctx);
    if (rc) return rc;
    if (mem == HVB_DEVICE)
    {
        st->dTasks = tasks;
        st->dOut = out;
        return HVB_OK;
    }
    cudaSetDevice(ctx->device);
    const size_t inPad = (inBytes + 255) & ~size_t(255);
    const size_t total = inPad + ((outBytes + 255) & ~size_t(255));
    if (ctx->pipelined && hvbIsPinned(tasks) && (!outBytes || hvbIsPinned(out)))
    {
        // a slot of the ring: task copy on the copy-in stream, kernels wait for it; hvbStageOut sends the results back
        // on the copy-out stream and the call returns without waiting
        const int id = ctx->nextSlot;
        ctx->nextSlot = (ctx->nextSlot + 1) % 4;
        hvb_context::Slot &slot = ctx->slots[id];
        cudaError_t e = cudaSuccess;
        if (slot.busy) e = cudaEventSynchronize(slot.done);
        slot.busy = false;
        if (e == cudaSuccess && slot.bytes < total)
        {
            if (slot.dev) e = cudaFree(slot.dev);
            slot.dev = nullptr;
            slot.bytes = 0;
            size_t cap = 1 << 20;
            while (cap < total) cap *= 2;
            if (e == cudaSuccess) e = cudaMalloc(&slot.dev, cap);
            if (e == cudaSuccess) slot.bytes = cap;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(slot.dev, tasks, inBytes, cudaMemcpyHostToDevice, ctx->copyIn);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->evIn, ctx->copyIn);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->evIn, 0);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "stage in (pipelined)");
        st->dTasks = slot.dev;
        st->dOut = static_cast<char *>(slot.dev) + inPad;
        st->slot = id;
        return HVB_OK;
    }
    if (ctx->hostStageBytes < total)
    {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->hostStage) cudaFreeHost(ctx->hostStage);
        if (ctx->devStage) cudaFree(ctx->devStage);
        ctx->hostStage = ctx->devStage = nullptr;
        ctx->hostStageBytes = ctx->devStageBytes = 0;
        size_t cap = 1 << 20;
        while (cap < total) cap *= 2;
        cudaError_t e = cudaMallocHost(&ctx->hostStage, cap);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->devStage, cap);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "staging buffers");
        ctx->hostStageBytes = ctx->devStageBytes = cap;
    }
    // Page-locked caller memory (cudaHostAlloc / cudaHostRegister, e.g. an encoder's task arena) is copied from
    // directly; pageable memory goes through the context's pinned staging buffer first.
    const void *hostSrc = tasks;
    if (!hvbIsPinned(tasks))
    {
        memcpy(ctx->hostStage, tasks, inBytes);
        hostSrc = ctx->hostStage;
    }
    cudaError_t e = cudaMemcpyAsync(ctx->devStage, hostSrc, inBytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "stage in");
    st->dTasks = ctx->devStage;
    st->dOut = static_cast<char *>(ctx->devStage) + inPad;
    st->hOutPinned = static_cast<char *>(ctx->hostStage) + inPad;
    return HVB_OK;
}

int hvbStageOut(hvb_context *ctx, void *out, size_t outBytes, hvb_mem mem, const HvbStaged &st)
{
    if (mem == HVB_DEVICE) return HVB_OK;
    cudaError_t e = cudaSuccess;
    if (st.slot >= 0)
    {
        hvb_context::Slot &slot = ctx->slots[st.slot];
        e = cudaEventRecord(ctx->evCompute, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyOut, ctx->evCompute, 0);
        if (e == cudaSuccess && outBytes) e = cudaMemcpyAsync(out, st.dOut, outBytes, cudaMemcpyDeviceToHost, ctx->copyOut);
        if (e == cudaSuccess) e = cudaEventRecord(slot.done, ctx->copyOut);
        slot.busy = true;
        return hvbCuda(ctx, e, "stage out (pipelined)");
    }
    const bool direct = outBytes && hvbIsPinned(out);
    if (outBytes) e = cudaMemcpyAsync(direct ? out : st.hOutPinned, st.dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "stage out");
    if (outBytes && !direct) memcpy(out, st.hOutPinned, outBytes);
    return HVB_OK;
}
