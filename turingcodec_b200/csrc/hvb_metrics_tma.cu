// hvb_metrics_tma.cu -- SAD / SAD4 with the blocks staged through shared memory by the Tensor Memory Accelerator.
//
// Same semantics as sadKernel / sad4Kernel of hvb_metrics.cu (havoc/sad.cpp:432-449, :513-542; 16-bit samples >> 2).
// What differs is how the bytes arrive.  A motion-search candidate starts at an arbitrary byte, so the load/store path
// has to fetch two aligned 16-byte chunks and funnel them for every 16 bytes it wants (twice the L1 wavefronts of an
// aligned block, profiles/r02a_stream.json: 63 % of the HBM roofline instead of 74 %, SAD4 50 % instead of 90 %).  The TMA
// takes the block's coordinates as they are: `cp.async.bulk.tensor.2d` drops a box of the picture plane into shared
// memory, aligned, signalling an mbarrier with the bytes it delivered; the lanes then read aligned 128-bit chunks.
//
// The TMA takes box coordinates whose innermost byte offset is a multiple of 16 only (measured on the B200: any other
// coordinate faults the kernel with "illegal instruction"), so a task with an operand off that grid -- the usual
// motion-search candidate -- is computed by the warp straight from global memory instead; the staged path serves
// co-located and 16-byte aligned operands, which is where it was measured against the load/store kernels.
//
// Tensor maps: one per plane and box shape, over the plane's whole allocation (padding included, so a vector into the
// padding is an ordinary coordinate), element type UINT8 for both sample widths (a 16-bit plane is a byte plane twice as
// wide).  Every box is 1 KB: 16 B x 64 rows, 32 B x 32 rows or 64 B x 16 rows, chosen per task by the block's width in
// bytes; a 128-byte-wide block (64 samples of 16 bits) is walked as two 64-byte halves.  A warp owns a ring of stages of
// (1 + NREF) boxes and its own mbarriers: lane 0 issues, all lanes consume, the ring keeps kStages - 1 items in flight
// across task boundaries.
#include "hvb_internal.cuh"

#include <cuda.h>

namespace {

constexpr int kWarps = 8;
constexpr int kBoxBytes = 1024;
constexpr int kShapes = 3; // box widths 16, 32, 64 bytes

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void barInit(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void barExpect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void barWait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\t"
                 "bra WAIT;\n\t"
                 "DONE:\n\t}" ::"r"(smemAddr(bar)),
                 "r"(parity)
                 : "memory");
}

__device__ __forceinline__ void tmaLoad2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smemAddr(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smemAddr(bar))
                 : "memory");
}

// a task of either kind, reduced to what the pipeline needs
struct Item
{
    int srcPlane, srcX, srcY;        // plane index (pic * 3 + cIdx), sample coordinates
    int refPlane, refX[4], refY[4];
    int w, h;
};

template <int NREF>
__device__ __forceinline__ Item loadItem(const void *tasks, int t)
{
    Item it;
    if (NREF == 4)
    {
        const hvb_sad4_task task = static_cast<const hvb_sad4_task *>(tasks)[t];
        it.srcPlane = task.src.pic * 3 + task.src.cIdx, it.srcX = task.src.x, it.srcY = task.src.y;
        it.refPlane = task.ref_pic * 3 + task.ref_cIdx;
#pragma unroll
        for (int k = 0; k < 4; ++k) it.refX[k] = task.rx[k], it.refY[k] = task.ry[k];
        it.w = task.w, it.h = task.h;
    }
    else
    {
        const hvb_metric_task task = static_cast<const hvb_metric_task *>(tasks)[t];
        it.srcPlane = task.a.pic * 3 + task.a.cIdx, it.srcX = task.a.x, it.srcY = task.a.y;
        it.refPlane = task.b.pic * 3 + task.b.cIdx, it.refX[0] = task.b.x, it.refY[0] = task.b.y;
        it.w = task.w, it.h = task.h;
    }
    return it;
}

// box shape of a block `wb` bytes wide: index, width in bytes, rows; 128-byte blocks are two halves of shape 2
__device__ __forceinline__ void shapeOf(int wb, int &shape, int &boxW, int &boxH, int &halves)
{
    shape = wb <= 16 ? 0 : (wb <= 32 ? 1 : 2);
    boxW = 16 << shape;
    boxH = kBoxBytes / boxW;
    halves = wb > 64 ? 2 : 1;
}

// every operand of the item starts on a 16-byte boundary of its plane's allocation
template <int NREF, int B>
__device__ __forceinline__ bool boxAligned(const HvbPlane *__restrict__ planes, const Item &item)
{
    unsigned bits = (unsigned)((planes[item.srcPlane].reserved + item.srcX) * B);
    const int reserved = planes[item.refPlane].reserved;
#pragma unroll
    for (int k = 0; k < NREF; ++k) bits |= (unsigned)((reserved + item.refX[k]) * B);
    return (bits & 15u) == 0;
}

template <typename Sample, int NREF, int STAGES>
__global__ void __launch_bounds__(kWarps * 32)
    sadTmaKernel(const HvbPlane *__restrict__ planes, const CUtensorMap *__restrict__ maps, const void *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    constexpr int B = (int)sizeof(Sample);
    constexpr int kStageBytes = (1 + NREF) * kBoxBytes;
    extern __shared__ __align__(128) uint8_t stageMem[]; // [warp][stage][1 + NREF boxes]
    __shared__ uint64_t bars[kWarps][STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *myStages = stageMem + (size_t)warp * STAGES * kStageBytes;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) barInit(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    const int warpsTotal = gridDim.x * kWarps;
    const int first = blockIdx.x * kWarps + warp;
    // issue cursor and consume cursor walk the same sequence of (task, strip, half)
    int it = first, iStrip = 0, iHalf = 0, issued = 0;
    int ct = first, cStrip = 0, cHalf = 0, consumed = 0;
    unsigned acc[NREF];
#pragma unroll
    for (int k = 0; k < NREF; ++k) acc[k] = 0;

    auto issueOne = [&]() {
        Item item;
        for (;; it += warpsTotal) // tasks with an operand off the 16-byte grid are not staged
        {
            if (it >= n) return false;
            item = loadItem<NREF>(tasks, it);
            if (boxAligned<NREF, B>(planes, item)) break;
        }
        int shape, boxW, boxH, halves;
        shapeOf(item.w * B, shape, boxW, boxH, halves);
        const int stage = issued % STAGES;
        if (lane == 0)
        {
            uint64_t *bar = &bars[warp][stage];
            uint8_t *dst = myStages + (size_t)stage * kStageBytes;
            barExpect(bar, (uint32_t)kStageBytes);
            const HvbPlane sp = planes[item.srcPlane], rp = planes[item.refPlane];
            const int dy = iStrip * boxH, dxBytes = iHalf * 64;
            tmaLoad2d(dst, maps + item.srcPlane * kShapes + shape, (sp.reserved + item.srcX) * B + dxBytes, sp.pad + item.srcY + dy, bar);
#pragma unroll
            for (int k = 0; k < NREF; ++k)
                tmaLoad2d(dst + (1 + k) * kBoxBytes, maps + item.refPlane * kShapes + shape, (rp.reserved + item.refX[k]) * B + dxBytes,
                          rp.pad + item.refY[k] + dy, bar);
        }
        ++issued;
        const int strips = (item.h + boxH - 1) / boxH;
        if (++iHalf == halves)
        {
            iHalf = 0;
            if (++iStrip == strips)
            {
                iStrip = 0;
                it += warpsTotal;
            }
        }
        return true;
    };

    for (int s = 0; s < STAGES - 1; ++s) issueOne();
    while (ct < n)
    {
        const Item item = loadItem<NREF>(tasks, ct);
        if (!boxAligned<NREF, B>(planes, item))
        {
            // straight from global memory, a lane per sample
            const HvbPlane sp = planes[item.srcPlane], rp = planes[item.refPlane];
            const Sample *src = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)item.srcY * sp.stride + item.srcX;
            for (int i = lane; i < item.w * item.h; i += 32)
            {
                const int y = i / item.w, x = i - y * item.w;
                const int a = src[(intptr_t)y * sp.stride + x];
#pragma unroll
                for (int k = 0; k < NREF; ++k)
                    acc[k] += (unsigned)abs(a - (int)(reinterpret_cast<const Sample *>(rp.base) + (intptr_t)(item.refY[k] + y) * rp.stride + item.refX[k])[x]);
            }
#pragma unroll
            for (int k = 0; k < NREF; ++k)
            {
                int v = hvbWarpSum((int)acc[k]);
                if (B == 2) v >>= 2;
                if (lane == 0) out[ct * NREF + k] = v;
                acc[k] = 0;
            }
            ct += warpsTotal;
            continue;
        }
        issueOne(); // keeps STAGES - 1 items in flight while this one is consumed (the stage it targets was freed last turn)
        int shape, boxW, boxH, halves;
        const int wb = item.w * B;
        shapeOf(wb, shape, boxW, boxH, halves);
        const int stage = consumed % STAGES;
        barWait(&bars[warp][stage], (uint32_t)((consumed / STAGES) & 1));
        const uint8_t *base = myStages + (size_t)stage * kStageBytes;
        const int rows = min(boxH, item.h - cStrip * boxH);       // rows of this strip that belong to the block
        const int rowBytes = min(boxW, wb - cHalf * 64);          // bytes of a row that belong to the block
        const int cpr = boxW >> 4;                                // 16-byte chunks per box row
#pragma unroll
        for (int j = 0; j < kBoxBytes / 16 / 32; ++j)
        {
            const int c = lane + 32 * j, row = c / cpr, col = (c - row * cpr) << 4;
            if (row < rows && col < rowBytes)
            {
                const int valid = rowBytes - col;
                uint4 vs = *reinterpret_cast<const uint4 *>(base + c * 16);
                if (valid < 16)
                {
                    const uint32_t m0 = valid >= 4 ? 0xffffffffu : (1u << (8 * valid)) - 1u;
                    const uint32_t m1 = valid >= 8 ? 0xffffffffu : (valid <= 4 ? 0u : (1u << (8 * (valid - 4))) - 1u);
                    const uint32_t m2 = valid >= 12 ? 0xffffffffu : (valid <= 8 ? 0u : (1u << (8 * (valid - 8))) - 1u);
                    const uint32_t m3 = valid <= 12 ? 0u : (1u << (8 * (valid - 12))) - 1u;
                    vs.x &= m0, vs.y &= m1, vs.z &= m2, vs.w &= m3;
#pragma unroll
                    for (int k = 0; k < NREF; ++k)
                    {
                        uint4 vr = *reinterpret_cast<const uint4 *>(base + (1 + k) * kBoxBytes + c * 16);
                        vr.x &= m0, vr.y &= m1, vr.z &= m2, vr.w &= m3;
                        acc[k] += B == 1 ? __vsadu4(vs.x, vr.x) + __vsadu4(vs.y, vr.y) + __vsadu4(vs.z, vr.z) + __vsadu4(vs.w, vr.w)
                                         : __vsadu2(vs.x, vr.x) + __vsadu2(vs.y, vr.y) + __vsadu2(vs.z, vr.z) + __vsadu2(vs.w, vr.w);
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < NREF; ++k)
                    {
                        const uint4 vr = *reinterpret_cast<const uint4 *>(base + (1 + k) * kBoxBytes + c * 16);
                        acc[k] += B == 1 ? __vsadu4(vs.x, vr.x) + __vsadu4(vs.y, vr.y) + __vsadu4(vs.z, vr.z) + __vsadu4(vs.w, vr.w)
                                         : __vsadu2(vs.x, vr.x) + __vsadu2(vs.y, vr.y) + __vsadu2(vs.z, vr.z) + __vsadu2(vs.w, vr.w);
                    }
                }
            }
        }
        __syncwarp(); // every lane has read its chunks: the stage may be refilled
        ++consumed;
        const int strips = (item.h + boxH - 1) / boxH;
        if (++cHalf == halves)
        {
            cHalf = 0;
            if (++cStrip == strips)
            {
#pragma unroll
                for (int k = 0; k < NREF; ++k)
                {
                    int v = hvbWarpSum((int)acc[k]);
                    if (B == 2) v >>= 2;
                    if (lane == 0) out[ct * NREF + k] = v;
                    acc[k] = 0;
                }
                cStrip = 0;
                ct += warpsTotal;
            }
        }
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encoder()
{
    static EncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiled>(p);
    }();
    return fn;
}

} // namespace

// (re)build the tensor maps of every live picture whose planes this context owns or imported; device copy at ctx->dTensorMaps
int hvbSyncTensorMaps(hvb_context *ctx)
{
    if (!ctx->tensorMapsDirty && ctx->dTensorMaps) return HVB_OK;
    EncodeTiled encode = encoder();
    if (!encode) return hvbFail(ctx, HVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cudaSetDevice(ctx->device);
    const size_t count = (size_t)HVB_MAX_PICTURES * 3 * kShapes;
    if (!ctx->dTensorMaps)
    {
        cudaError_t e = cudaMalloc(&ctx->dTensorMaps, count * sizeof(CUtensorMap));
        if (e != cudaSuccess) return hvbCuda(ctx, e, "tensor maps");
    }
    std::vector<CUtensorMap> host(count);
    memset(host.data(), 0, count * sizeof(CUtensorMap));
    for (int pic = 0; pic < HVB_MAX_PICTURES; ++pic)
    {
        const HvbPicture &p = ctx->pictures[pic];
        if (!p.live) continue;
        for (int c = 0; c < 3; ++c)
        {
            const HvbPlane &pl = p.plane[c];
            if (!pl.base || !p.tmaBase[c]) continue; // wrapped pictures have no device allocation to map
            const cuuint64_t strideBytes = (cuuint64_t)pl.stride * ctx->bps;
            const cuuint64_t dims[2] = {strideBytes, (cuuint64_t)p.tmaRows[c]};
            const cuuint64_t strides[1] = {strideBytes};
            const cuuint32_t elem[2] = {1, 1};
            for (int s = 0; s < kShapes; ++s)
            {
                const cuuint32_t box[2] = {(cuuint32_t)(16 << s), (cuuint32_t)(kBoxBytes / (16 << s))};
                const CUresult r = encode(&host[((size_t)pic * 3 + c) * kShapes + s], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p.tmaBase[c], dims, strides, box, elem,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return hvbFail(ctx, HVB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
            }
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->dTensorMaps, host.data(), count * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "tensor maps");
    ctx->tensorMapsDirty = false;
    return HVB_OK;
}

template <typename Sample, int NREF>
static int launch(hvb_context *ctx, const void *dTasks, int n, int32_t *dOut)
{
    constexpr int STAGES = NREF == 4 ? 3 : 4;
    const size_t smem = (size_t)kWarps * STAGES * (1 + NREF) * kBoxBytes;
    auto kernel = sadTmaKernel<Sample, NREF, STAGES>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "sadTmaKernel attributes");
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kWarps * 32, smem);
    if (perSm < 1) perSm = 1;
    int blocks = (n + kWarps - 1) / kWarps;
    if (blocks > ctx->smCount * perSm) blocks = ctx->smCount * perSm;
    kernel<<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, static_cast<const CUtensorMap *>(ctx->dTensorMaps), dTasks, n, dOut);
    HVB_LAUNCH_CHECK(ctx, "sadTmaKernel");
    return HVB_OK;
}

// nref: 1 (hvb_metric_task, SAD) or 4 (hvb_sad4_task)
int hvbLaunchSadTma(hvb_context *ctx, const void *dTasks, int n, int32_t *dOut, int nref)
{
    int rc = hvbSyncTensorMaps(ctx);
    if (rc) return rc;
    if (ctx->bps == 1) return nref == 4 ? launch<uint8_t, 4>(ctx, dTasks, n, dOut) : launch<uint8_t, 1>(ctx, dTasks, n, dOut);
    return nref == 4 ? launch<uint16_t, 4>(ctx, dTasks, n, dOut) : launch<uint16_t, 1>(ctx, dTasks, n, dOut);
}
