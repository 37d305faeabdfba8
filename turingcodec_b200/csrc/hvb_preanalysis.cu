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
//   AdaptiveQuantisation::preAnalysis                       turing/AdaptiveQuantisation.h:172-246
//   ShotChangeDetection: luma histogram, getLikelihood      turing/SCDetection.h:71-147, :237-262
//
// Adaptive-quantisation activity: per layer the plane is cut into units, each unit into four quadrants; a thread block
// takes a 64 x 64 region of a layer (one unit of the coarsest layer, 256 of the finest), its threads walk the region in
// runs of four samples and add each run's sum and sum of squares to the (unit, quadrant) accumulators in shared memory
// (integer sums: order-free); one thread per unit then forms the four integer variances with the reference's quotients
// and quirks.  Algorithmic bytes: the luma plane once per layer (wh B), 8 bytes out per unit.
//
// Shot-change detection: the 64-bin histogram is a shared-memory histogram per block flushed with one atomic per bin; the
// block statistics of getLikelihood are double-precision sums whose order of addition is the reference's, so a thread
// walks a whole block serially (52 blocks per picture pair; the pass is called for the rare picture pair whose histogram
// difference falls between the two thresholds).  Every double operation is a single correctly rounded IEEE operation
// (no contraction into FMA), which is what the reference's x86-64 code executes.
//
// All kernels here use no warp-level primitive: tests/test_host_emulated_preanalysis.py runs their source on the CPU.
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


// sample -> the 8-bit value ShotChangeDetection works on (turing/SCDetection.h:244, :284: 16-bit samples >> 2, truncated to a byte)
__device__ __forceinline__ int scdByte(uint8_t v) { return v; }
__device__ __forceinline__ int scdByte(uint16_t v) { return (v >> 2) & 0xff; }

// single IEEE operations: nvcc contracts a * b + c into an FMA unless told otherwise; the host emulation's g++ (x86-64, no
// -mfma) does not contract
__device__ __forceinline__ double addRn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
__device__ __forceinline__ double mulRn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__device__ __forceinline__ double divRn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

constexpr int kAqRegion = 64;                    // a block's share of a layer: 64 x 64 samples
constexpr int kAqMaxUnits = (kAqRegion / 4) * (kAqRegion / 4); // units of 4 x 4

template <typename Sample>
__global__ void __launch_bounds__(256)
    aqActivityKernel(const HvbPlane *__restrict__ planes, const hvb_aq_layer_task *__restrict__ tasks, int n, long long *__restrict__ out)
{
    // per (unit of the region, quadrant): sum, sum of squares
    __shared__ unsigned long long acc[kAqMaxUnits][4][2];
    // jobs: the 64 x 64 regions of every task, task-major
    int jobBase = 0;
    for (int ti = 0; ti < n; ++ti)
    {
        const hvb_aq_layer_task t = tasks[ti];
        const HvbPlane &pl = planes[t.pic * 3];
        const int W = pl.width, H = pl.height, U = t.unit;
        const int regionsX = (W + kAqRegion - 1) / kAqRegion, regionsY = (H + kAqRegion - 1) / kAqRegion;
        const int unitsPerRow = (W + U - 1) / U;
        const int upr = kAqRegion / U; // units per region row
        const Sample *base = reinterpret_cast<const Sample *>(pl.base);
        for (int job = blockIdx.x; job < regionsX * regionsY; job += gridDim.x)
        {
            const int ry = job / regionsX, rx = job - ry * regionsX;
            const int x0 = rx * kAqRegion, y0 = ry * kAqRegion;
            const int rw = min(kAqRegion, W - x0), rh = min(kAqRegion, H - y0);
            for (int i = threadIdx.x; i < upr * upr * 8; i += blockDim.x) (&acc[0][0][0])[i] = 0;
            __syncthreads();
            // runs of four samples: a run lies in one row; its samples are filed one by one when the run straddles a boundary
            const int runsPerRow = (rw + 3) >> 2;
            for (int i = threadIdx.x; i < runsPerRow * rh; i += blockDim.x)
            {
                const int y = i / runsPerRow, xr = (i - y * runsPerRow) << 2;
                const Sample *row = base + (intptr_t)(y0 + y) * pl.stride + x0;
                const int uy = y / U, ly = y - uy * U;
                const int uh = min(U, H - (y0 + uy * U));
                const int qy = ly >= (uh >> 1) ? 2 : 0;
                int key = -1;
                unsigned long long s = 0, q = 0;
                for (int k = 0; k < 4 && xr + k < rw; ++k)
                {
                    const int x = xr + k, ux = x / U, lx = x - ux * U;
                    const int uw = min(U, W - (x0 + ux * U));
                    const int now = ((uy * upr + ux) << 2) | qy | (lx >= (uw >> 1) ? 1 : 0);
                    if (now != key)
                    {
                        if (key >= 0)
                        {
                            atomicAdd(&acc[key >> 2][key & 3][0], s);
                            atomicAdd(&acc[key >> 2][key & 3][1], q);
                        }
                        key = now, s = 0, q = 0;
                    }
                    const unsigned long long v = row[x];
                    s += v, q += v * v;
                }
                if (key >= 0)
                {
                    atomicAdd(&acc[key >> 2][key & 3][0], s);
                    atomicAdd(&acc[key >> 2][key & 3][1], q);
                }
            }
            __syncthreads();
            for (int u = threadIdx.x; u < upr * upr; u += blockDim.x)
            {
                const int uy = u / upr, ux = u - uy * upr;
                const int col = x0 + ux * U, rowPic = y0 + uy * U;
                if (col >= W || rowPic >= H) continue;
                const int uw = min(U, W - col), uh = min(U, H - rowPic);
                const int num = (uw * uh) >> 2;
                long long minVar = 0;
                if (num)
                {
                    unsigned long long sum[4], sq[4];
                    for (int b = 0; b < 4; ++b) sum[b] = acc[u][b][0], sq[b] = acc[u][b][1];
                    // the reference's quadrant 0 keeps the square of the last sample it visits, quadrant 1 adds the samples themselves
                    sq[0] = 0;
                    if ((uw >> 1) && (uh >> 1))
                    {
                        const unsigned long long v = base[(intptr_t)(rowPic + (uh >> 1) - 1) * pl.stride + col + (uw >> 1) - 1];
                        sq[0] = v * v;
                    }
                    sq[1] = sum[1];
                    for (int b = 0; b < 4; ++b)
                    {
                        const long long average = (long long)(sum[b] / (unsigned long long)num);
                        const long long variance = (long long)(sq[b] / (unsigned long long)num) - average * average;
                        if (b == 0 || variance < minVar) minVar = variance;
                    }
                }
                out[t.out + (rowPic / U) * unitsPerRow + col / U] = minVar;
            }
            __syncthreads();
        }
        jobBase += regionsX * regionsY;
    }
    (void)jobBase;
}

template <typename Sample>
__global__ void __launch_bounds__(256)
    scdHistogramKernel(const HvbPlane *__restrict__ planes, const int16_t *__restrict__ pics, int n, int *__restrict__ out)
{
    __shared__ int hist[64];
    for (int pi = 0; pi < n; ++pi)
    {
        const HvbPlane &pl = planes[pics[pi] * 3];
        const Sample *base = reinterpret_cast<const Sample *>(pl.base);
        const int chunksPerRow = (pl.width + 15) >> 4; // a thread takes 16 consecutive samples of a row
        const int chunks = chunksPerRow * pl.height;
        const int perBlock = (chunks + gridDim.x - 1) / gridDim.x;
        const int first = blockIdx.x * perBlock, last = min(chunks, first + perBlock);
        if (first >= last) continue;
        for (int i = threadIdx.x; i < 64; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int c = first + threadIdx.x; c < last; c += blockDim.x)
        {
            const int y = c / chunksPerRow, x = (c - y * chunksPerRow) << 4;
            const Sample *row = base + (intptr_t)y * pl.stride + x;
            const int count = min(16, pl.width - x);
            int bin = -1, run = 0;
            for (int k = 0; k < count; ++k)
            {
                const int b = scdByte(row[k]) >> 2;
                if (b != bin)
                {
                    if (run) atomicAdd(&hist[bin], run);
                    bin = b, run = 0;
                }
                ++run;
            }
            if (run) atomicAdd(&hist[bin], run);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64; i += blockDim.x)
            if (hist[i]) atomicAdd(&out[64 * pi + i], hist[i]);
        __syncthreads();
    }
}

// blocks of a task's grid: rows j = margin*bh, .. < height - margin*bh, columns likewise
__host__ __device__ inline int scdGridCount(int size, int block, int margin)
{
    int count = 0;
    if (block > 0)
        for (int j = margin * block; j < size - margin * block; j += block) ++count;
    return count;
}

template <typename Sample>
__global__ void __launch_bounds__(64)
    scdBlockStatsKernel(const HvbPlane *__restrict__ planes, const hvb_scd_stats_task *__restrict__ tasks, int n, double *__restrict__ out)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    int jobBase = 0;
    for (int ti = 0; ti < n; ++ti)
    {
        const hvb_scd_stats_task t = tasks[ti];
        const HvbPlane &pl = planes[t.pic * 3];
        const Sample *base = reinterpret_cast<const Sample *>(pl.base);
        const int W = pl.width, H = pl.height, bw = W >> 3, bh = H >> 3;
        const int cols = scdGridCount(W, bw, t.margin), rows = scdGridCount(H, bh, t.margin);
        // thread (jobBase + k) mod gthreads takes block k, so that the blocks of all tasks spread over the grid
        for (int k = (gtid - jobBase % gthreads + gthreads) % gthreads; k < rows * cols; k += gthreads)
        {
            const int r = k / cols, c = k - r * cols;
            const long long origin = (long long)(t.margin + r) * bh * W + (long long)(t.margin + c) * bw;
            // element (h, w) is byte origin + h * HEIGHT + w of the packed plane (turing/SCDetection.h:87)
            long long sum = 0;
            for (int h = 0; h < bh; ++h)
                for (int w = 0; w < bw; ++w)
                {
                    const long long f = origin + (long long)h * H + w;
                    const int y = (int)(f / W), x = (int)(f - (long long)y * W);
                    sum += scdByte(base[(intptr_t)y * pl.stride + x]);
                }
            const double count = (double)(bh * bw);
            const double avg = divRn((double)sum, count);
            double var = 0.0;
            for (int h = 0; h < bh; ++h)
                for (int w = 0; w < bw; ++w)
                {
                    const long long f = origin + (long long)h * H + w;
                    const int y = (int)(f / W), x = (int)(f - (long long)y * W);
                    const double d = addRn((double)scdByte(base[(intptr_t)y * pl.stride + x]), -avg);
                    var = addRn(var, mulRn(d, d));
                }
            out[t.out + 2 * k] = avg;
            out[t.out + 2 * k + 1] = divRn(var, count);
        }
        jobBase += rows * cols;
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


extern "C" int hvb_aq_activity_batch(hvb_context *ctx, const hvb_aq_layer_task *tasks, int n, int64_t *out, int outCount, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && outCount >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    if (mem == HVB_HOST)
        for (int i = 0; i < n; ++i)
        {
            const hvb_aq_layer_task &t = tasks[i];
            HVB_CHECK_ARGS(ctx, t.pic >= 0 && t.pic < HVB_MAX_PICTURES && ctx->pictures[t.pic].live);
            HVB_CHECK_ARGS(ctx, t.unit >= 4 && t.unit <= 64 && !(t.unit & (t.unit - 1)));
            const HvbPicture &p = ctx->pictures[t.pic];
            const long long units = (long long)((p.width + t.unit - 1) / t.unit) * ((p.height + t.unit - 1) / t.unit);
            HVB_CHECK_ARGS(ctx, t.out >= 0 && t.out + units <= outCount);
        }
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int64_t) * outCount, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_aq_layer_task *>(st.dTasks);
    auto *dO = static_cast<long long *>(st.dOut);
    const int blocks = ctx->smCount * 8;
    if (ctx->bps == 1)
        aqActivityKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    else
        aqActivityKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    HVB_LAUNCH_CHECK(ctx, "aqActivityKernel");
    return hvbStageOut(ctx, out, sizeof(int64_t) * outCount, mem, st);
}

extern "C" int hvb_scd_histogram_batch(hvb_context *ctx, const int16_t *pics, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (pics && out)));
    if (!n) return HVB_OK;
    if (mem == HVB_HOST)
        for (int i = 0; i < n; ++i) HVB_CHECK_ARGS(ctx, pics[i] >= 0 && pics[i] < HVB_MAX_PICTURES && ctx->pictures[pics[i]].live);
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, pics, sizeof(int16_t) * n, out, sizeof(int32_t) * 64 * n, mem, &st);
    if (rc) return rc;
    const auto *dP = static_cast<const int16_t *>(st.dTasks);
    auto *dO = static_cast<int *>(st.dOut);
    cudaError_t e = cudaMemsetAsync(dO, 0, sizeof(int32_t) * 64 * n, ctx->stream); // the blocks add their partial histograms
    if (e != cudaSuccess) return hvbCuda(ctx, e, "hvb_scd_histogram_batch");
    const int blocks = ctx->smCount * 4;
    if (ctx->bps == 1)
        scdHistogramKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dP, n, dO);
    else
        scdHistogramKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dP, n, dO);
    HVB_LAUNCH_CHECK(ctx, "scdHistogramKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 64 * n, mem, st);
}

extern "C" int hvb_scd_block_stats_batch(hvb_context *ctx, const hvb_scd_stats_task *tasks, int n, double *out, int outCount, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && outCount >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    int total = 0;
    if (mem == HVB_HOST)
        for (int i = 0; i < n; ++i)
        {
            const hvb_scd_stats_task &t = tasks[i];
            HVB_CHECK_ARGS(ctx, t.pic >= 0 && t.pic < HVB_MAX_PICTURES && ctx->pictures[t.pic].live && t.margin >= 0 && t.margin <= 3);
            const HvbPicture &p = ctx->pictures[t.pic];
            const int bw = p.width >> 3, bh = p.height >> 3;
            HVB_CHECK_ARGS(ctx, bw > 0 && bh > 0);
            const int cols = scdGridCount(p.width, bw, t.margin), rows = scdGridCount(p.height, bh, t.margin);
            HVB_CHECK_ARGS(ctx, t.out >= 0 && t.out + 2 * rows * cols <= outCount);
            if (rows && cols)
            {
                // the last element the reference's addressing reaches must lie inside its width x height vector
                const long long last = (long long)(t.margin + rows - 1) * bh * p.width + (long long)(t.margin + cols - 1) * bw +
                                       (long long)(bh - 1) * p.height + bw - 1;
                if (last >= (long long)p.width * p.height)
                    return hvbFail(ctx, HVB_ERR_INVALID, "hvb_scd_block_stats_batch: the reference's block addressing leaves the plane for this picture size");
            }
            total += rows * cols;
        }
    else
        total = outCount / 2;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(double) * outCount, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_scd_stats_task *>(st.dTasks);
    auto *dO = static_cast<double *>(st.dOut);
    const int blocks = max(1, min(ctx->smCount, total)); // a thread per block of the grid, one thread per thread block: each gets an SM's L1 to itself
    if (ctx->bps == 1)
        scdBlockStatsKernel<uint8_t><<<blocks, 1, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    else
        scdBlockStatsKernel<uint16_t><<<blocks, 1, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO);
    HVB_LAUNCH_CHECK(ctx, "scdBlockStatsKernel");
    return hvbStageOut(ctx, out, sizeof(double) * outCount, mem, st);
}
