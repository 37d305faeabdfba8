// hvb_metrics.cu -- batched SAD / SAD4 / SSD / Hadamard-SATD over device-resident pictures.
//
// Reference semantics (bit-exact):
//   havoc_sad           havoc/sad.cpp:432-449     (u16: >> 2)
//   havoc_sad_multiref  havoc/sad.cpp:513-542     (4 references, u16: each >> 2)
//   havoc_ssd           havoc/ssd.cpp:28-43       (uint32 wrap, u16: >> 4)
//   havoc_hadamard_satd havoc/hadamard.cpp:58-98  (2x2 / 4x4 (s+1)>>1 / 8x8 (s+2)>>2, u16: >> 2)
//   measureSatd         turing/Measure.h:96-135   (tiling of a w x h block)
//
// Mapping: one warp per candidate.  These kernels are pure streaming reductions (every sample is
// read once per candidate), so the bound is memory: 2*w*h*B algorithmic bytes per candidate
// (5*w*h*B for SAD4).  Source rows are 4-byte aligned by construction (PU x0 is a multiple of 4,
// plane rows are 256-byte aligned); reference rows are arbitrarily aligned and are read as two
// aligned words + funnel shift.  Byte lanes are reduced with the SIMD-in-word video instructions
// (VABSDIFF4 / dp4a), then a 5-step shuffle tree.
#include "hvb_internal.cuh"
#include "hvb_satd.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

template <typename Sample>
__device__ __forceinline__ int sadBlock(const Sample *a, int sa, const Sample *b, int sb, int w, int h, int lane)
{
    int acc = 0;
    if (sizeof(Sample) == 1 && !(w & 15) &&
        !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | (uintptr_t)sa | (uintptr_t)sb) & 15))
    {
        // streaming case (both blocks 16-byte aligned, e.g. co-located blocks): one 128-bit load per operand per lane
        // per step -- this is the path the HBM roofline is quoted on (tools/stream_metrics.py)
        const int cpr = w >> 4, total = cpr * h; // 16-byte chunks per row / in the block
        const uint8_t *pa = reinterpret_cast<const uint8_t *>(a), *pb = reinterpret_cast<const uint8_t *>(b);
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / cpr, x = (i - y * cpr) << 4;
            const uint4 va = __ldg(reinterpret_cast<const uint4 *>(pa + y * sa + x)), vb = __ldg(reinterpret_cast<const uint4 *>(pb + y * sb + x));
            acc = __vsadu4(va.x, vb.x) + __vsadu4(va.y, vb.y) + __vsadu4(va.z, vb.z) + __vsadu4(va.w, vb.w) + acc;
        }
    }
    else if (sizeof(Sample) == 1 && !(w & 3))
    {
        const int wq = w >> 2, total = wq * h;
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / wq, x = (i - y * wq) << 2;
            const uint32_t va = hvbLoad4u8(reinterpret_cast<const uint8_t *>(a) + y * sa + x);
            const uint32_t vb = hvbLoad4u8(reinterpret_cast<const uint8_t *>(b) + y * sb + x);
            acc = __vsadu4(va, vb) + acc;
        }
    }
    else
    {
        const int total = w * h;
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / w, x = i - y * w;
            acc += abs((int)a[y * sa + x] - (int)b[y * sb + x]);
        }
    }
    return acc;
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sadKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    for (int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_metric_task task = tasks[t];
        int sa, sb;
        const Sample *a = hvbBlockPtr<Sample>(planes, task.a, sa);
        const Sample *b = hvbBlockPtr<Sample>(planes, task.b, sb);
        int acc = hvbWarpSum(sadBlock<Sample>(a, sa, b, sb, task.w, task.h, lane));
        if (sizeof(Sample) == 2) acc >>= 2;
        if (lane == 0) out[t] = acc;
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sad4Kernel(const HvbPlane *__restrict__ planes, const hvb_sad4_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    for (int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_sad4_task task = tasks[t];
        int ss;
        const Sample *src = hvbBlockPtr<Sample>(planes, task.src, ss);
        const HvbPlane &rp = planes[task.ref_pic * 3 + task.ref_cIdx];
        const Sample *rbase = reinterpret_cast<const Sample *>(rp.base);
        const int sr = rp.stride;
        const int w = task.w, h = task.h;
        int acc[4] = {0, 0, 0, 0};
        const Sample *ref[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ref[k] = rbase + (intptr_t)task.ry[k] * sr + task.rx[k];

        if (sizeof(Sample) == 1 && !(w & 3))
        {
            const int wq = w >> 2, total = wq * h;
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / wq, x = (i - y * wq) << 2;
                const uint32_t vs = hvbLoad4u8(reinterpret_cast<const uint8_t *>(src) + y * ss + x);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[k] = __vsadu4(vs, hvbLoad4u8(reinterpret_cast<const uint8_t *>(ref[k]) + y * sr + x)) + acc[k];
            }
        }
        else
        {
            const int total = w * h;
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / w, x = i - y * w;
                const int s = src[y * ss + x];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] += abs(s - (int)ref[k][y * sr + x]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            int v = hvbWarpSum(acc[k]);
            if (sizeof(Sample) == 2) v >>= 2;
            if (lane == 0) out[t * 4 + k] = v;
        }
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    ssdKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    for (int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_metric_task task = tasks[t];
        int sa, sb;
        const Sample *a = hvbBlockPtr<Sample>(planes, task.a, sa);
        const Sample *b = hvbBlockPtr<Sample>(planes, task.b, sb);
        const int w = task.w, h = task.h;
        unsigned acc = 0;
        if (sizeof(Sample) == 1 && !(w & 15) &&
            !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | (uintptr_t)sa | (uintptr_t)sb) & 15))
        {
            const int cpr = w >> 4, total = cpr * h;
            const uint8_t *pa = reinterpret_cast<const uint8_t *>(a), *pb = reinterpret_cast<const uint8_t *>(b);
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / cpr, x = (i - y * cpr) << 4;
                const uint4 va = __ldg(reinterpret_cast<const uint4 *>(pa + y * sa + x)), vb = __ldg(reinterpret_cast<const uint4 *>(pb + y * sb + x));
                uint32_t d;
                d = __vabsdiffu4(va.x, vb.x); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(va.y, vb.y); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(va.z, vb.z); acc = __dp4a(d, d, acc);
                d = __vabsdiffu4(va.w, vb.w); acc = __dp4a(d, d, acc);
            }
        }
        else if (sizeof(Sample) == 1 && !(w & 3))
        {
            const int wq = w >> 2, total = wq * h;
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / wq, x = (i - y * wq) << 2;
                const uint32_t va = hvbLoad4u8(reinterpret_cast<const uint8_t *>(a) + y * sa + x);
                const uint32_t vb = hvbLoad4u8(reinterpret_cast<const uint8_t *>(b) + y * sb + x);
                const uint32_t d = __vabsdiffu4(va, vb);
                acc = __dp4a(d, d, acc);
            }
        }
        else
        {
            const int total = w * h;
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / w, x = i - y * w;
                const int d = (int)a[y * sa + x] - (int)b[y * sb + x];
                acc += (unsigned)(d * d);
            }
        }
        acc = hvbWarpSumU(acc);
        if (sizeof(Sample) == 2) acc >>= 4;
        if (lane == 0) out[t] = acc;
    }
}

} // namespace

namespace {

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    satdKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    for (int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_metric_task task = tasks[t];
        int sa, sb;
        const Sample *a = hvbBlockPtr<Sample>(planes, task.a, sa);
        const Sample *b = hvbBlockPtr<Sample>(planes, task.b, sb);
        // one register-resident Hadamard tile per lane.  (A row-per-lane variant with the vertical butterfly in
        // shuffles was measured at 0.93 vs 1.48 TB/s on 32x32 blocks and dropped: 27 shuffles per tile row cost more
        // than the idle lanes; at ~700 instructions per 128 bytes the kernel sits on the issue roofline near 50% of HBM.)
        int acc = hvbMeasureSatdLanes<Sample, Sample>(a, sa, b, sb, task.w, task.h, lane, 32, sizeof(Sample) == 2 ? 2 : 0);
        acc = hvbWarpSum(acc);
        if (lane == 0) out[t] = acc;
    }
}

template <typename Task>
int gridFor(hvb_context *ctx, int n)
{
    const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int cap = ctx->smCount * 8; // 8 resident 256-thread CTAs per SM
    return blocks < cap ? blocks : cap;
}

} // namespace

#define HVB_DISPATCH_SAMPLE(ctx, kernel, grid, ...)                                                      \
    do                                                                                                   \
    {                                                                                                    \
        if ((ctx)->bps == 1)                                                                             \
            kernel<uint8_t><<<(grid), kWarpsPerBlock * 32, 0, (ctx)->stream>>>(__VA_ARGS__);             \
        else                                                                                             \
            kernel<uint16_t><<<(grid), kWarpsPerBlock * 32, 0, (ctx)->stream>>>(__VA_ARGS__);           \
    } while (0)

extern "C" int hvb_sad_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    HVB_DISPATCH_SAMPLE(ctx, sadKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "sadKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
}

extern "C" int hvb_sad4_batch(hvb_context *ctx, const hvb_sad4_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * 4 * n, mem, &st);
    if (rc) return rc;
    HVB_DISPATCH_SAMPLE(ctx, sad4Kernel, gridFor<hvb_sad4_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_sad4_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "sad4Kernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 4 * n, mem, st);
}

extern "C" int hvb_ssd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, uint32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(uint32_t) * n, mem, &st);
    if (rc) return rc;
    HVB_DISPATCH_SAMPLE(ctx, ssdKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<uint32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "ssdKernel");
    return hvbStageOut(ctx, out, sizeof(uint32_t) * n, mem, st);
}

extern "C" int hvb_satd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    HVB_DISPATCH_SAMPLE(ctx, satdKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "satdKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
}
