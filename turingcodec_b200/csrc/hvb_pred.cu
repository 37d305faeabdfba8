// hvb_pred.cu -- batched HEVC inter prediction (uni / bi), SubtractBi and the stand-alone fused
// interpolation + SATD (the sub-pel refinement of the search is hvb_me_subpel.cu, the PU cost hvb_pu_cost.cu).
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
