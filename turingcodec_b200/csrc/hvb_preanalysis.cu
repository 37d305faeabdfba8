// hvb_preanalysis.cu -- per-picture pre-analysis passes on device-resident source pictures.
//
// Reference semantics (bit-exact):
//   EstimateIntraComplexity::computeSatd8x8 / preAnalysis   turing/EstimateIntraComplexity.h:55-176
//
// Intra complexity: the AC Hadamard energy of every 8x8 luma block of the source (the rate control's measure of how
// expensive an intra picture will be).  A thread per block: eight 8-sample rows as vector loads (consecutive threads take
// consecutive blocks of a block row, so a warp reads whole 256-byte row segments), the transform in registers, one int out.
// Algorithmic bytes: the luma plane once (wh B) plus 4 bytes per 64 samples.
//
// Status: written after the round's GPU budget was spent; bit-exact under host emulation
// (tests/test_host_emulated_preanalysis.py); tests/test_gpu_zz_preanalysis.py has not yet run on a GPU.
#include "hvb_internal.cuh"

namespace {

// eight consecutive samples at an 8-sample-aligned address
__device__ __forceinline__ void load8(const uint8_t *p, int (&v)[8])
{
    const uint2 w = *reinterpret_cast<const uint2 *>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        v[i] = (w.x >> (8 * i)) & 0xff;
        v[4 + i] = (w.y >> (8 * i)) & 0xff;
    }
}
__device__ __forceinline__ void load8(const uint16_t *p, int (&v)[8])
{
    const uint4 w = *reinterpret_cast<const uint4 *>(p);
    v[0] = w.x & 0xffff, v[1] = w.x >> 16, v[2] = w.y & 0xffff, v[3] = w.y >> 16;
    v[4] = w.z & 0xffff, v[5] = w.z >> 16, v[6] = w.w & 0xffff, v[7] = w.w >> 16;
}

// in-place 8-point Hadamard butterflies over m[base + k * step], k = 0..7 (compile-time indices once unrolled)
template <int STEP>
__device__ __forceinline__ void hadamard8(int (&m)[64], int base)
{
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1)
#pragma unroll
        for (int b = 0; b < 8; b += 2 * half)
#pragma unroll
            for (int j = 0; j < half; ++j)
            {
                const int p = m[base + (b + j) * STEP], q = m[base + (b + j + half) * STEP];
                m[base + (b + j) * STEP] = p + q;
                m[base + (b + j + half) * STEP] = p - q;
            }
}

template <typename Sample>
__global__ void __launch_bounds__(128)
    intraComplexityKernel(const HvbPlane *__restrict__ planes, const hvb_intra_complexity_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int ti = 0; ti < n; ++ti)
    {
        const hvb_intra_complexity_task t = tasks[ti];
        const HvbPlane &pl = planes[t.pic * 3];
        const Sample *base = reinterpret_cast<const Sample *>(pl.base) + (intptr_t)t.y0 * pl.stride + t.x0;
        const int blocks = t.wBlocks * t.hBlocks;
        for (int b = gtid; b < blocks; b += gthreads)
        {
            const int by = b / t.wBlocks, bx = b - by * t.wBlocks;
            const Sample *p = base + (intptr_t)(8 * by) * pl.stride + 8 * bx;
            int m[64];
#pragma unroll
            for (int y = 0; y < 8; ++y)
            {
                int row[8];
                load8(p + (intptr_t)y * pl.stride, row);
#pragma unroll
                for (int x = 0; x < 8; ++x) m[8 * y + x] = row[x];
            }
#pragma unroll
            for (int y = 0; y < 8; ++y) hadamard8<1>(m, 8 * y);
#pragma unroll
            for (int x = 0; x < 8; ++x) hadamard8<8>(m, x);
            int total = 0;
#pragma unroll
            for (int i = 1; i < 64; ++i) total += abs(m[i]); // without the DC term m[0]
            total = (total + 2) >> 2;
            if (sizeof(Sample) == 2) total >>= 2;
            out[t.out + b] = total;
        }
    }
}

} // namespace

extern "C" int hvb_intra_complexity_batch(hvb_context *ctx, const hvb_intra_complexity_task *tasks, int n, int32_t *out, int outCount, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && outCount >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * outCount, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_intra_complexity_task *>(st.dTasks);
    auto *dO = static_cast<int32_t *>(st.dOut);
    const int blocks = ctx->smCount * 8;
    if (ctx->bps == 1)
        intraComplexityKernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    else
        intraComplexityKernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    HVB_LAUNCH_CHECK(ctx, "intraComplexityKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * outCount, mem, st);
}
