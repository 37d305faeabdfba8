// hvb_internal.cuh -- shared host/device definitions of libhvb (not part of the public ABI).
#pragma once

#include "../../include/hvb.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

// ---------------------------------------------------------------------------------------------
// Device-side view of a picture plane.  `base` points at sample (0,0); samples at negative
// coordinates down to -pad and beyond width/height up to +pad are valid memory (edge padding).
// Rows are `stride` samples apart; stride*bps is a multiple of 256 bytes and base-pad*bps is
// 256-byte aligned, so TMA descriptors and 128-bit loads on row starts are legal.
// ---------------------------------------------------------------------------------------------
struct HvbPlane
{
    void *base;
    int32_t stride; // in samples
    int32_t width, height;
    int32_t pad;
    int32_t reserved; // column of sample 0 in an allocated row: sample (x, y) sits at column reserved + x of row pad + y (tensor-map coordinates)
};

static const int HVB_MAX_PICTURES = 1024;

struct HvbPicture
{
    bool live = false;
    int width = 0, height = 0, pad = 0;
    void *alloc[3] = {nullptr, nullptr, nullptr};
    size_t allocBytes[3] = {0, 0, 0};
    HvbPlane plane[3];
    void *tmaBase[3] = {nullptr, nullptr, nullptr}; // start of each plane's device allocation and its rows (tensor maps span the
    int tmaRows[3] = {0, 0, 0};                     // whole allocation, padding included); kept by imports, null for wrapped pictures
    void *lfInfo = nullptr; // deblocking side information (hvb_deblock_info_upload): block records, then CTU records
    size_t lfBytes = 0;
    void *saoInfo = nullptr; // SAO records per CTU (hvb_sao_info_upload)
    size_t saoBytes = 0;
};

// device-side view of a picture's deblocking side information (hvb_loopfilter.cu)
struct HvbLoopInfo
{
    const hvb_deblock_block *blocks;
    const hvb_deblock_ctu *ctus;
    const hvb_sao_ctu *sao;
    int32_t blockStride, blockRows, widthInCtbs, ctbLog2, saoCount, reserved;
};

struct hvb_context
{
    int device = 0;
    int bps = 1;
    int bitDepth = 8;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t ownStream = nullptr;
    int64_t launches = 0;
    cudaEvent_t marks[16] = {}; // hvb_mark
    std::string lastError;

    HvbPicture pictures[HVB_MAX_PICTURES];
    HvbPlane *dPlanes = nullptr; // [HVB_MAX_PICTURES*3] mirrored on device
    bool planesDirty = true;
    void *dTensorMaps = nullptr; // [HVB_MAX_PICTURES * 3 * 3] CUtensorMap (hvb_metrics_tma.cu), built on first use
    bool tensorMapsDirty = true;
    bool useTma = false; // hvb_set_tma
    int tuFusedMax = 256; // hvb_set_tu_fused_max: batches of at most this many transform blocks take the one-launch form of hvb_tu_chain_batch
    HvbLoopInfo *dLoopInfo = nullptr; // [HVB_MAX_PICTURES] on the device, allocated by the first hvb_deblock_info_upload
    HvbLoopInfo loopInfoHost[HVB_MAX_PICTURES] = {};
    int *workCursors = nullptr; // [64] device-side task cursors of the persistent kernels (zeroed on the stream before each use)

    // pipelined host mode (hvb_set_pipelined): copies on their own streams, a ring of staging slots
    bool pipelined = false;
    cudaStream_t copyIn = nullptr, copyOut = nullptr;
    cudaEvent_t evIn = nullptr, evCompute = nullptr; // stream-to-stream dependencies (re-recorded per use)
    struct Slot
    {
        void *dev = nullptr;
        size_t bytes = 0;
        cudaEvent_t done = nullptr; // the call that used the slot has delivered its results
        bool busy = false;
    } slots[4];
    int nextSlot = 0;

    // staging for HVB_HOST calls
    void *hostStage = nullptr;  // pinned
    size_t hostStageBytes = 0;
    void *devStage = nullptr;
    size_t devStageBytes = 0;

    // pools
    void *samplePool = nullptr; // neighbour samples (bps each)
    size_t samplePoolCount = 0;
    void *slab = nullptr; // hvb_picture_reserve: one allocation the following hvb_picture_create calls carve their planes from
    size_t slabBytes = 0, slabUsed = 0;
    int16_t *coeffPool = nullptr;
    size_t coeffPoolCount = 0;
    hvb_rdoq_ctx *rdoqCtx = nullptr;
    int rdoqCtxCount = 0;
    // the caller's page-locked arrays in place of the two above (hvb_coeff_pool_wrap, hvb_rdoq_contexts_wrap): one-launch TU chain only
    int16_t *coeffWrap = nullptr;
    size_t coeffWrapCount = 0;
    const hvb_rdoq_ctx *rdoqWrap = nullptr;
    int rdoqWrapCount = 0;
    int2 *rdoqBits = nullptr; // [rdoqCtxCount * sizeof(hvb_rdoq_ctx)] bit costs of both bins per context state byte
    size_t rdoqBitsCount = 0;
    int *rdoqLast = nullptr; // [rdoqCtxCount * 160] last-position prefix rates per snapshot (hvb_rdoq.cuh lastPrefixRate)
    size_t rdoqLastCount = 0;
    void *scratch = nullptr; // kernel workspace (RDOQ per-TU state, ...)
    size_t scratchBytes = 0;
};

// --- host helpers (hvb_context.cu) -----------------------------------------------------------
int hvbFail(hvb_context *ctx, int status, const char *what);
int hvbCuda(hvb_context *ctx, cudaError_t e, const char *what);
int hvbSyncPlanes(hvb_context *ctx);
int hvbLaunchSadTma(hvb_context *ctx, const void *dTasks, int n, int32_t *dOut, int nref); // hvb_metrics_tma.cu
bool hvbEnvTma();                                                                          // HVB_TMA=1 in the environment: default of hvb_set_tma
int hvbEnsureScratch(hvb_context *ctx, size_t bytes);
int hvbEnsureCoeffPool(hvb_context *ctx, size_t count);
int hvbEnsureSamplePool(hvb_context *ctx, size_t count);
int hvbLaunchRdoqBits(hvb_context *ctx, int first, int count); // hvb_tu.cu
// Enqueue a host->device copy of caller memory.  Pipelined mode + page-locked source: on the copy-in stream (ordered after
// earlier uploads, not after earlier batches); the compute stream waits for it before anything issued later; no host wait.
// Otherwise on the compute stream, followed by a host wait (the source may be pageable).
int hvbUpload(hvb_context *ctx, void *dev, size_t devPitch, const void *host, size_t hostPitch, size_t widthBytes, size_t rows,
              const char *what);
bool hvbIsPinned(const void *p);

// Stage `inBytes` of tasks to the device when mem == HVB_HOST and reserve `outBytes` of device
// result space behind them.  Returns device pointers for both.
struct HvbStaged
{
    const void *dTasks = nullptr;
    void *dOut = nullptr;
    void *hOutPinned = nullptr;
    int slot = -1; // >= 0: pipelined call, results are delivered asynchronously
};
int hvbStageIn(hvb_context *ctx, const void *tasks, size_t inBytes, void *out, size_t outBytes,
               hvb_mem mem, HvbStaged *st);
int hvbStageOut(hvb_context *ctx, void *out, size_t outBytes, hvb_mem mem, const HvbStaged &st);

#define HVB_CHECK_ARGS(ctx, cond)                                                   \
    do                                                                              \
    {                                                                               \
        if (!(ctx)) return HVB_ERR_INVALID;                                         \
        if (!(cond)) return hvbFail((ctx), HVB_ERR_INVALID, "invalid argument: " #cond); \
    } while (0)

#define HVB_LAUNCH_CHECK(ctx, name)                                 \
    do                                                              \
    {                                                               \
        (ctx)->launches++;                                          \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return hvbCuda((ctx), e__, name);   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

template <typename Sample>
__device__ __forceinline__ const Sample *hvbBlockPtr(const HvbPlane *planes, const hvb_block &b, int &stride)
{
    const HvbPlane &p = planes[b.pic * 3 + b.cIdx];
    stride = p.stride;
    return reinterpret_cast<const Sample *>(p.base) + (intptr_t)b.y * p.stride + b.x;
}

template <typename Sample>
__device__ __forceinline__ Sample *hvbBlockPtrW(const HvbPlane *planes, const hvb_block &b, int &stride)
{
    const HvbPlane &p = planes[b.pic * 3 + b.cIdx];
    stride = p.stride;
    return reinterpret_cast<Sample *>(p.base) + (intptr_t)b.y * p.stride + b.x;
}

__device__ __forceinline__ int hvbWarpSum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned hvbWarpSumU(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int hvbClip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// Unaligned 4 x u8 load: two aligned 32-bit loads and a byte funnel.
__device__ __forceinline__ uint32_t hvbLoad4u8(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3);
    const uint32_t lo = __ldg(q);
    if (sh == 0) return lo;
    const uint32_t hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, sh * 8);
}

#endif // __CUDACC__
