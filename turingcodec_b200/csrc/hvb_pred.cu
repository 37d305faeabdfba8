// hvb_pred.cu -- batched HEVC inter prediction (uni / bi), SubtractBi and the fused
// interpolation + SATD used by sub-pel motion refinement.
//
// Reference semantics (bit-exact):
//   HavocPredUni   havoc/pred_inter.cpp:76-202   copy / h / v / hv, 8-tap luma, 4-tap chroma
//   HavocPredBi    havoc/pred_inter.cpp:1207-1252 two 14-bit predictions, rounded mean, clip
//   SubtractBi     havoc/pred_inter.cpp:2063-2080 clip(2*src - pred)
//   costDistortionMv (distortion part)  turing/Search.hpp:1965-1981 + Measure.h:96-135
//
// One separable formulation covers the reference's four uni cases exactly:
//   mid = (sum_k cx[k] * s[x+k-m])              >> shift1        (shift1 = min(4, bd-8))
//   out = clip((sum_k cy[k] * mid[y+k-m] + rnd) >> (6 + shift3)) (shift3 = max(2, 14-bd))
// because a zero fraction selects the {..,64,..} kernel, for which the pass is an exact left shift
// (see DESIGN.md "interpolation identity"); intermediates fit int16 for 8..10-bit input.
//
// Mapping: one warp per PU.  The horizontal pass writes a (h+taps-1) x w int16 tile into the warp's
// private shared-memory slice, the vertical pass reads it back column-wise (conflict-free: lanes
// walk x).  The fused kernel keeps the predicted block in shared memory and runs register-resident
// Hadamard tiles against the source block, so a sub-pel candidate costs (w+7)(h+7)B + whB bytes of
// reads and 4 bytes of writes, never a round trip of the prediction through HBM.
#include "hvb_internal.cuh"
#include "hvb_satd.cuh"
#include "hvb_interp.cuh"

namespace {

using namespace hvb_interp;
constexpr int kWarps = 4;

template <typename Sample, int TAPS>
__device__ void predictWarp(int16_t *smem, const HvbPlane *planes, const hvb_pred_task &t, int cIdx, int bitDepth, int lane)
{
    int16_t *mid = smem, *first = smem + kMidElems;
    const int w = t.w, h = t.h;
    const int fracMask = TAPS == 8 ? 3 : 7, fracShift = TAPS == 8 ? 2 : 3;
    const int shift1 = min(4, bitDepth - 8), shift3 = max(2, 14 - bitDepth);
    const bool bi = t.ref_pic[1] >= 0;
    int sd;
    Sample *dst = hvbBlockPtrW<Sample>(planes, t.dst, sd);
    const int maxv = (1 << bitDepth) - 1;

    for (int r = 0; r < (bi ? 2 : 1); ++r)
    {
        const HvbPlane &rp = planes[t.ref_pic[r] * 3 + cIdx];
        const int xFrac = t.mvx[r] & fracMask, yFrac = t.mvy[r] & fracMask;
        const Sample *ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)(t.y + (t.mvy[r] >> fracShift)) * rp.stride +
                            (t.x + (t.mvx[r] >> fracShift));
        passH<Sample, TAPS>(mid, ref, rp.stride, w, h, xFrac, shift1, lane);
        __syncwarp();
        int cy[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) cy[k] = coef<TAPS>(yFrac, k);
        const int total = w * h;
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / w, x = i - y * w;
            const int v = passV<TAPS>(mid, w, x, y, cy);
            if (!bi)
                dst[y * sd + x] = (Sample)hvbClip3(0, maxv, (v + (1 << (5 + shift3))) >> (6 + shift3));
            else if (r == 0)
                first[i] = (int16_t)(v >> 6);
            else
                dst[y * sd + x] = (Sample)hvbClip3(0, maxv, ((int)first[i] + (v >> 6) + (1 << shift3)) >> (shift3 + 1));
        }
        __syncwarp();
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    predKernel(const HvbPlane *__restrict__ planes, const hvb_pred_task *__restrict__ tasks, int n, int bitDepth)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int16_t *smem = reinterpret_cast<int16_t *>(smemRaw + warp * kSmemPerWarp);
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_pred_task t = tasks[i];
        const int cIdx = t.dst.cIdx;
        if (cIdx == 0)
            predictWarp<Sample, 8>(smem, planes, t, 0, bitDepth, lane);
        else
            predictWarp<Sample, 4>(smem, planes, t, cIdx, bitDepth, lane);
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    interpSatdKernel(const HvbPlane *__restrict__ planes, const hvb_interp_satd_task *__restrict__ tasks, int n,
                     int32_t *__restrict__ out, int bitDepth)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int16_t *mid = reinterpret_cast<int16_t *>(smemRaw + warp * kSmemPerWarp);
    int16_t *pred = mid + kMidElems;
    const int warpsTotal = gridDim.x * kWarps;
    const int shift1 = min(4, bitDepth - 8), shift3 = max(2, 14 - bitDepth);
    const int maxv = (1 << bitDepth) - 1;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_interp_satd_task t = tasks[i];
        const int w = t.w, h = t.h;
        int ss;
        const Sample *src = hvbBlockPtr<Sample>(planes, t.src, ss);
        const HvbPlane &rp = planes[t.ref_pic * 3];
        const Sample *ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)(t.src.y + (t.mvy >> 2)) * rp.stride +
                            (t.src.x + (t.mvx >> 2));
        passH<Sample, 8>(mid, ref, rp.stride, w, h, t.mvx & 3, shift1, lane);
        __syncwarp();
        int cy[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cy[k] = coef<8>(t.mvy & 3, k);
        const int total = w * h;
        for (int j = lane; j < total; j += 32)
        {
            const int y = j / w, x = j - y * w;
            pred[j] = (int16_t)hvbClip3(0, maxv, (passV<8>(mid, w, x, y, cy) + (1 << (5 + shift3))) >> (6 + shift3));
        }
        __syncwarp();
        int acc = hvbMeasureSatdLanes<Sample, int16_t>(src, ss, pred, w, w, h, lane, 32, sizeof(Sample) == 2 ? 2 : 0);
        acc = hvbWarpSum(acc);
        if (lane == 0) out[i] = acc;
        __syncwarp();
    }
}

template <typename Sample>
__global__ void __launch_bounds__(256)
    subtractBiKernel(const HvbPlane *__restrict__ planes, const hvb_subtract_bi_task *__restrict__ tasks, int n, int bitDepth)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * 8;
    const int maxv = (1 << bitDepth) - 1;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += warpsTotal)
    {
        const hvb_subtract_bi_task t = tasks[i];
        int sd, sp, ss;
        Sample *dst = hvbBlockPtrW<Sample>(planes, t.dst, sd);
        const Sample *pred = hvbBlockPtr<Sample>(planes, t.pred, sp);
        const Sample *src = hvbBlockPtr<Sample>(planes, t.src, ss);
        const int w = t.w, total = t.w * t.h;
        for (int j = lane; j < total; j += 32)
        {
            const int y = j / w, x = j - y * w;
            dst[y * sd + x] = (Sample)hvbClip3(0, maxv, 2 * (int)src[y * ss + x] - (int)pred[y * sp + x]);
        }
    }
}

// ---- measurePuCost's distortion: predictInter + SATD of Y, Cb, Cr (turing/Search.hpp:1668-1682) ----------------
// predictUni / predictBi (turing/Dsp.h:769-864): the luma block origin is xPb + (mv >> 2) clamped by
// clipMvLumaComponent (:723-731); the chroma origin is that value >> 1 and the chroma phase is mv & 7.

__device__ __forceinline__ int clipMvLumaComponent(int component, int nPbSize, int pictureSize)
{
    if (component + nPbSize + 4 < 0) return -nPbSize - 4;
    if (component > pictureSize + 2) return pictureSize + 2;
    return component;
}

// one colour component of one PU: prediction into the warp's shared tile (optionally stored), SATD against the source
template <typename Sample, int TAPS>
__device__ int puComponent(int16_t *smem, const HvbPlane *planes, const hvb_pu_cost_task &t, int cIdx, const int (&bx)[2],
                           const int (&by)[2], int bitDepth, int lane)
{
    int16_t *mid = smem, *pred = smem + kMidElems;
    const int sh = cIdx ? 1 : 0, w = t.w >> sh, h = t.h >> sh, total = w * h;
    const int fracMask = TAPS == 8 ? 3 : 7;
    const int shift1 = min(4, bitDepth - 8), shift3 = max(2, 14 - bitDepth);
    const int maxv = (1 << bitDepth) - 1;
    const bool bi = t.ref_pic[0] >= 0 && t.ref_pic[1] >= 0;
    bool first = true;
    for (int r = 0; r < 2; ++r)
    {
        if (t.ref_pic[r] < 0) continue;
        const HvbPlane &rp = planes[t.ref_pic[r] * 3 + cIdx];
        const Sample *ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)(by[r] >> sh) * rp.stride + (bx[r] >> sh);
        passH<Sample, TAPS>(mid, ref, rp.stride, w, h, t.mvx[r] & fracMask, shift1, lane);
        __syncwarp();
        int cy[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) cy[k] = coef<TAPS>(t.mvy[r] & fracMask, k);
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / w, x = i - y * w;
            const int v = passV<TAPS>(mid, w, x, y, cy);
            if (!bi)
                pred[i] = (int16_t)hvbClip3(0, maxv, (v + (1 << (5 + shift3))) >> (6 + shift3));
            else if (first)
                pred[i] = (int16_t)(v >> 6);
            else
                pred[i] = (int16_t)hvbClip3(0, maxv, ((int)pred[i] + (v >> 6) + (1 << shift3)) >> (shift3 + 1));
        }
        __syncwarp();
        first = false;
    }
    if (t.dst_pic >= 0)
    {
        const HvbPlane &dp = planes[t.dst_pic * 3 + cIdx];
        Sample *dst = reinterpret_cast<Sample *>(dp.base) + (intptr_t)(t.y0 >> sh) * dp.stride + (t.x0 >> sh);
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / w, x = i - y * w;
            dst[y * dp.stride + x] = (Sample)pred[i];
        }
    }
    if (cIdx && ((w | h) & 3)) return 0; // Compute<Satd, Rectangle> (turing/Measure.h:156-160)
    const HvbPlane &sp = planes[t.src_pic * 3 + cIdx];
    const Sample *src = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)(t.y0 >> sh) * sp.stride + (t.x0 >> sh);
    const int acc = hvbMeasureSatdLanes<Sample, int16_t>(src, sp.stride, pred, w, w, h, lane, 32, sizeof(Sample) == 2 ? 2 : 0);
    return hvbWarpSum(acc);
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    puCostKernel(const HvbPlane *__restrict__ planes, const hvb_pu_cost_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
                 int bitDepth)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int16_t *smem = reinterpret_cast<int16_t *>(smemRaw + warp * kSmemPerWarp);
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_pu_cost_task t = tasks[i];
        int bx[2] = {0, 0}, by[2] = {0, 0};
        for (int r = 0; r < 2; ++r)
            if (t.ref_pic[r] >= 0)
            {
                const HvbPlane &rp = planes[t.ref_pic[r] * 3];
                bx[r] = clipMvLumaComponent(t.x0 + (t.mvx[r] >> 2), t.w, rp.width);
                by[r] = clipMvLumaComponent(t.y0 + (t.mvy[r] >> 2), t.h, rp.height);
            }
        const int y = puComponent<Sample, 8>(smem, planes, t, 0, bx, by, bitDepth, lane);
        __syncwarp();
        const int cb = puComponent<Sample, 4>(smem, planes, t, 1, bx, by, bitDepth, lane);
        __syncwarp();
        const int cr = puComponent<Sample, 4>(smem, planes, t, 2, bx, by, bitDepth, lane);
        __syncwarp();
        if (lane == 0)
        {
            out[3 * i] = y;
            out[3 * i + 1] = cb;
            out[3 * i + 2] = cr;
        }
    }
}

int gridWarps(hvb_context *ctx, int n, int warps, int perSm)
{
    const int blocks = (n + warps - 1) / warps;
    const int cap = ctx->smCount * perSm;
    return blocks < cap ? blocks : cap;
}

} // namespace

extern "C" int hvb_pred_batch(hvb_context *ctx, const hvb_pred_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const int smem = kWarps * kSmemPerWarp;
    const auto *dT = static_cast<const hvb_pred_task *>(st.dTasks);
    if (ctx->bps == 1)
    {
        cudaFuncSetAttribute(predKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        predKernel<uint8_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, ctx->bitDepth);
    }
    else
    {
        cudaFuncSetAttribute(predKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        predKernel<uint16_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "predKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_interp_satd_batch(hvb_context *ctx, const hvb_interp_satd_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    const int smem = kWarps * kSmemPerWarp;
    const auto *dT = static_cast<const hvb_interp_satd_task *>(st.dTasks);
    auto *dO = static_cast<int32_t *>(st.dOut);
    if (ctx->bps == 1)
    {
        cudaFuncSetAttribute(interpSatdKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        interpSatdKernel<uint8_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    else
    {
        cudaFuncSetAttribute(interpSatdKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        interpSatdKernel<uint16_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "interpSatdKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
}

extern "C" int hvb_subtract_bi_batch(hvb_context *ctx, const hvb_subtract_bi_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_subtract_bi_task *>(st.dTasks);
    if (ctx->bps == 1)
        subtractBiKernel<uint8_t><<<gridWarps(ctx, n, 8, 8), 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, ctx->bitDepth);
    else
        subtractBiKernel<uint16_t><<<gridWarps(ctx, n, 8, 8), 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "subtractBiKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

int hvbPuCostBatchV1(hvb_context *ctx, const hvb_pu_cost_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * 3 * n, mem, &st);
    if (rc) return rc;
    const int smem = kWarps * kSmemPerWarp;
    const auto *dT = static_cast<const hvb_pu_cost_task *>(st.dTasks);
    auto *dO = static_cast<int32_t *>(st.dOut);
    if (ctx->bps == 1)
    {
        cudaFuncSetAttribute(puCostKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        puCostKernel<uint8_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    else
    {
        cudaFuncSetAttribute(puCostKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        puCostKernel<uint16_t><<<gridWarps(ctx, n, kWarps, 3), kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "puCostKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 3 * n, mem, st);
}
