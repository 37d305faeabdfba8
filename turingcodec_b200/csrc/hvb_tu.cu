// hvb_tu.cu -- batched transform / quantisation primitives and the fused TU pipeline.
//
// Reference semantics (bit-exact):
//   havoc::Transform (forward DCT-II 4..32, DST-VII 4)   havoc/transform.cpp:3071-3397
//        two passes, shifts log2n-1+bd-8 and log2n+6, each result truncated to int16 WITH WRAP (:3071-3084)
//   havoc::inverse_transform(_add)                        havoc/transform.cpp:50-401, transform.h:104-114
//        two passes, shifts 7 and 20-bd, each clipped to int16; add to prediction, clip to bit depth
//   havoc_quantize / havoc_quantize_inverse               havoc/quantize.cpp:278-304, :37-46
//   TU pipeline                                           turing/Reconstruct.cpp:731-857 (inter), :180-356 (intra)
//
// The partial butterflies of the reference are a factorisation of plain integer matrix products
// (C = M X M^T); integer sums are associative and nothing overflows int32, so the products are
// evaluated directly: one warp per transform block, lanes across the output frequency (forward) or
// output sample (inverse) index so that the matrix row is read conflict-free from shared memory and
// the block operand is a broadcast.  The fused pipeline keeps residual and coefficients in the warp's
// shared-memory slice: per TU it reads src and pred (2 n^2 B), writes rec (n^2 B), the levels (2 n^2) and
// 16 bytes of results.  RDOQ blocks additionally pass through a workspace between the front stage and
// the serial stage: the unquantised coefficients (2 B each) and one 32-byte record per coefficient, laid
// out by a per-batch compaction (scanLocalKernel / scanBlocksKernel below), so the workspace is sized by
// the RDOQ blocks of the batch, not by the capacity of the coefficient pool.
#include "hvb_internal.cuh"
#include "hvb_rdoq.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kBlk = 32 * 32;

// HEVC core transform: M32[k][i], generated from its 31 distinct magnitudes (see oracle_havoc.c dct_coeff).
__device__ __constant__ int8_t kMag[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                           61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
__device__ __constant__ int8_t kDst[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

struct alignas(16) Matrices
{
    int8_t m[32][32];  // m[k][i]
    int8_t mt[32][32]; // mt[i][k] = m[k][i]
    int8_t dst[4][4];  // dst[k][i]
    int8_t dstT[4][4];
    int8_t m16[16][16];  // the 16-point matrix (rows 0, 2, 4, .. of m, first 16 columns) and its transpose, compact:
    int8_t m16t[16][16]; // A-fragment rows of the tensor-core passes must be contiguous
};

__device__ void initMatrices(Matrices &M)
{
    for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x)
    {
        const int k = idx >> 5, i = idx & 31;
        int v;
        if (k == 0)
            v = 64;
        else
        {
            int a = (k * (2 * i + 1)) & 127;
            if (a > 64) a = 128 - a;
            v = a > 32 ? -kMag[64 - a] : kMag[a];
        }
        M.m[k][i] = (int8_t)v;
        M.mt[i][k] = (int8_t)v;
        if (!(k & 1) && i < 16)
        {
            M.m16[k >> 1][i] = (int8_t)v;
            M.m16t[i][k >> 1] = (int8_t)v;
        }
    }
    if (threadIdx.x < 16)
    {
        const int k = threadIdx.x >> 2, i = threadIdx.x & 3;
        M.dst[k][i] = kDst[k][i];
        M.dstT[i][k] = kDst[k][i];
    }
}

// forward pass: out[k*n + j] = (int16)((sum_i M[k][i] * in[j*is + i] + add) >> shift)   (wraps)
__device__ __forceinline__ void fwdPass(const Matrices &M, int16_t *out, const int16_t *in, int is, int log2n, bool dst, int shift,
                                        int lane)
{
    const int n = 1 << log2n, step = 5 - log2n; // row k of the n-point matrix is row k << step of M32
    const int add = 1 << (shift - 1);
    const int k = lane & (n - 1), jj = lane >> log2n, jstep = 32 >> log2n;
    if (n == 32)
    {
        for (int j = 0; j < 32; ++j)
        {
            int acc = add;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) acc += (int)M.mt[i][k] * (int)in[j * is + i];
            out[k * 32 + j] = (int16_t)(acc >> shift);
        }
        return;
    }
    for (int j = jj; j < n; j += jstep)
    {
        int acc = add;
        for (int i = 0; i < n; ++i)
        {
            const int c = dst ? M.dstT[i][k] : M.mt[i][k << step];
            acc += c * (int)in[j * is + i];
        }
        out[k * n + j] = (int16_t)(acc >> shift);
    }
}

// ---- 16- and 32-point passes on the integer tensor cores ----------------------------------------------------------
// out[k][j] = f((sum_i A[k][i] * in[j][i] + add) >> shift), all three N x N row-major; f wraps to int16 (forward
// transform, transform.cpp:3071-3084) or clips (inverse).  A is a transform matrix (int8); the 16-bit operand is split
// into a low (unsigned) and a high (signed) byte plane on the fly, in = 256 hi + lo, so a pass is two IMMA m16n8k32
// products per 16x8 output tile -- s8 x u8 and s8 x s8 -- recombined in the epilogue (exact: integer sums).
//   forward  pass p:  A = M,   in = previous result              (fwdPass semantics: out[k][j] = sum_i M[k][i] in[j][i])
//   inverse  pass 1:  A = M^T, in = dequantised levels TRANSPOSED -> out = reference's first-pass result transposed
//   inverse  pass 2:  A = M^T, in = that, TRANSPOSE_OUT          -> the residual in natural layout
__device__ __forceinline__ void immaS8U8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void immaS8S8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int N, bool CLIP, bool TRANSPOSE_OUT>
__device__ __forceinline__ void mmaPass(const int8_t *A, int16_t *out, const int16_t *in, int shift, int lane)
{
    const int g = lane >> 2, t = lane & 3;
    const int add = 1 << (shift - 1);
    // A fragments: rows mt*16 + g (+8), columns 4t.. (+16)
    uint32_t a[N / 16][4];
#pragma unroll
    for (int mt = 0; mt < N / 16; ++mt)
    {
        const int8_t *r0 = A + (mt * 16 + g) * N + 4 * t;
        a[mt][0] = *reinterpret_cast<const uint32_t *>(r0);
        a[mt][1] = *reinterpret_cast<const uint32_t *>(r0 + 8 * N);
        a[mt][2] = N == 32 ? *reinterpret_cast<const uint32_t *>(r0 + 16) : 0u;
        a[mt][3] = N == 32 ? *reinterpret_cast<const uint32_t *>(r0 + 8 * N + 16) : 0u;
    }
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt)
    {
        // B fragments of column j = nt*8 + g: in[j][4t .. 4t+3] (and [16 + 4t ..]) as byte planes
        const int16_t *row = in + (nt * 8 + g) * N + 4 * t;
        const uint2 w0 = *reinterpret_cast<const uint2 *>(row);
        const uint32_t lo0 = __byte_perm(w0.x, w0.y, 0x6420), hi0 = __byte_perm(w0.x, w0.y, 0x7531);
        uint32_t lo1 = 0, hi1 = 0;
        if (N == 32)
        {
            const uint2 w1 = *reinterpret_cast<const uint2 *>(row + 16);
            lo1 = __byte_perm(w1.x, w1.y, 0x6420);
            hi1 = __byte_perm(w1.x, w1.y, 0x7531);
        }
#pragma unroll
        for (int mt = 0; mt < N / 16; ++mt)
        {
            int cl[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
            immaS8U8(cl, a[mt][0], a[mt][1], a[mt][2], a[mt][3], lo0, lo1);
            immaS8S8(ch, a[mt][0], a[mt][1], a[mt][2], a[mt][3], hi0, hi1);
            int v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int x = (cl[r] + (ch[r] << 8) + add) >> shift;
                v[r] = CLIP ? hvbClip3(-32768, 32767, x) : (int)(int16_t)x;
            }
            const int k = mt * 16 + g, j = nt * 8 + 2 * t;
            if (TRANSPOSE_OUT)
            {
                out[j * N + k] = (int16_t)v[0];
                out[(j + 1) * N + k] = (int16_t)v[1];
                out[j * N + k + 8] = (int16_t)v[2];
                out[(j + 1) * N + k + 8] = (int16_t)v[3];
            }
            else
            {
                *reinterpret_cast<uint32_t *>(out + k * N + j) = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16);
                *reinterpret_cast<uint32_t *>(out + (k + 8) * N + j) = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
            }
        }
    }
}

// forward transform of the block in sA (through sB) -- tensor cores for the 16- and 32-point DCT, matrix products in
// shared memory otherwise
__device__ __forceinline__ void forwardTransform(const Matrices &M, int16_t *sA, int16_t *sB, int log2n, bool dst, int bitDepth, int lane)
{
    const int nn = 1 << log2n, shift1 = log2n - 1 + bitDepth - 8, shift2 = log2n + 6;
    if (log2n == 5)
    {
        mmaPass<32, false, false>(&M.m[0][0], sB, sA, shift1, lane);
        __syncwarp();
        mmaPass<32, false, false>(&M.m[0][0], sA, sB, shift2, lane);
    }
    else if (log2n == 4)
    {
        mmaPass<16, false, false>(&M.m16[0][0], sB, sA, shift1, lane);
        __syncwarp();
        mmaPass<16, false, false>(&M.m16[0][0], sA, sB, shift2, lane);
    }
    else
    {
        fwdPass(M, sB, sA, nn, log2n, dst, shift1, lane);
        __syncwarp();
        fwdPass(M, sA, sB, nn, log2n, dst, shift2, lane);
    }
}

// does inverseTransform want its input transposed (in[j][i] = coefficient (row i, column j))?
__device__ __forceinline__ bool inverseWantsTransposed(int log2n) { return log2n >= 4; }

// inverse transform of the coefficients in sA (through sB); the residual ends in sA in natural layout
__device__ __forceinline__ void inverseTransform(const Matrices &M, int16_t *sA, int16_t *sB, int log2n, bool dst, int bitDepth, int lane);

// inverse pass: out[j*n + k] = clip16((sum_i M[i][k] * in[i*n + j] + add) >> shift)
__device__ __forceinline__ void invPass(const Matrices &M, int16_t *out, const int16_t *in, int log2n, bool dst, int shift, int lane)
{
    const int n = 1 << log2n, step = 5 - log2n;
    const int add = 1 << (shift - 1);
    const int k = lane & (n - 1), jj = lane >> log2n, jstep = 32 >> log2n;
    for (int j = jj; j < n; j += jstep)
    {
        int acc = add;
        for (int i = 0; i < n; ++i)
        {
            const int c = dst ? M.dst[i][k] : M.m[i << step][k];
            acc += c * (int)in[i * n + j];
        }
        out[j * n + k] = (int16_t)hvbClip3(-32768, 32767, acc >> shift);
    }
}

__device__ __forceinline__ void inverseTransform(const Matrices &M, int16_t *sA, int16_t *sB, int log2n, bool dst, int bitDepth, int lane)
{
    if (log2n == 5)
    {
        mmaPass<32, true, false>(&M.mt[0][0], sB, sA, 7, lane);
        __syncwarp();
        mmaPass<32, true, true>(&M.mt[0][0], sA, sB, 20 - bitDepth, lane);
    }
    else if (log2n == 4)
    {
        mmaPass<16, true, false>(&M.m16t[0][0], sB, sA, 7, lane);
        __syncwarp();
        mmaPass<16, true, true>(&M.m16t[0][0], sA, sB, 20 - bitDepth, lane);
    }
    else
    {
        invPass(M, sB, sA, log2n, dst, 7, lane);
        __syncwarp();
        invPass(M, sA, sB, log2n, dst, 20 - bitDepth, lane);
    }
}

__device__ __forceinline__ int quantOne(int v, int scale, int shift, int off)
{
    const int mag = (abs(v) * scale + off) >> shift;
    return hvbClip3(-32768, 32767, v < 0 ? -mag : mag);
}

__device__ __forceinline__ int dequantOne(int v, int scale, int shift)
{
    return hvbClip3(-32768, 32767, (v * scale + (1 << (shift - 1))) >> shift);
}

// ---- stand-alone primitives over the coefficient pool ---------------------------------------

__global__ void __launch_bounds__(kWarps * 32)
    transformKernel(int16_t *__restrict__ pool, const hvb_transform_task *__restrict__ tasks, int n, int bitDepth, int inverse)
{
    __shared__ Matrices M;
    __shared__ __align__(16) int16_t sA[kWarps][kBlk];
    __shared__ __align__(16) int16_t sB[kWarps][kBlk];
    initMatrices(M);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * kWarps;
    for (int t = blockIdx.x * kWarps + warp; t < n; t += warpsTotal)
    {
        const hvb_transform_task task = tasks[t];
        const int log2n = task.log2n, nn = 1 << log2n, count = nn * nn;
        const bool dst = task.trType != 0;
        if (!inverse)
        {
            for (int i = lane; i < count; i += 32) sA[warp][i] = pool[task.src + (i >> log2n) * task.src_stride + (i & (nn - 1))];
            __syncwarp();
            forwardTransform(M, sA[warp], sB[warp], log2n, dst, bitDepth, lane);
        }
        else
        {
            const bool tr = inverseWantsTransposed(log2n) && !dst;
            for (int i = lane; i < count; i += 32) sA[warp][tr ? ((i & (nn - 1)) << log2n) + (i >> log2n) : i] = pool[task.src + i];
            __syncwarp();
            inverseTransform(M, sA[warp], sB[warp], log2n, dst, bitDepth, lane);
        }
        __syncwarp();
        for (int i = lane; i < count; i += 32) pool[task.dst + i] = sA[warp][i];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
    quantKernel(int16_t *__restrict__ pool, const hvb_quant_task *__restrict__ tasks, int n, int32_t *__restrict__ cbf, int inverse)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * 8;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_quant_task task = tasks[t];
        int any = 0;
        if (!inverse)
        {
            const int off = task.offset << (task.shift - 16);
            for (int i = lane; i < task.n; i += 32)
            {
                const int q = quantOne(pool[task.src + i], task.scale, task.shift, off);
                any |= q;
                pool[task.dst + i] = (int16_t)q;
            }
            any = __any_sync(0xffffffffu, any != 0);
            if (lane == 0) cbf[t] = any;
        }
        else
            for (int i = lane; i < task.n; i += 32) pool[task.dst + i] = (int16_t)dequantOne(pool[task.src + i], task.scale, task.shift);
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    itaKernel(const HvbPlane *__restrict__ planes, const int16_t *__restrict__ pool, const hvb_ita_task *__restrict__ tasks, int n,
              int bitDepth)
{
    __shared__ Matrices M;
    __shared__ __align__(16) int16_t sA[kWarps][kBlk];
    __shared__ __align__(16) int16_t sB[kWarps][kBlk];
    initMatrices(M);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * kWarps;
    const int maxv = (1 << bitDepth) - 1;
    for (int t = blockIdx.x * kWarps + warp; t < n; t += warpsTotal)
    {
        const hvb_ita_task task = tasks[t];
        const int log2n = task.log2n, nn = 1 << log2n, count = nn * nn;
        const bool tr = inverseWantsTransposed(log2n) && task.trType == 0;
        for (int i = lane; i < count; i += 32) sA[warp][tr ? ((i & (nn - 1)) << log2n) + (i >> log2n) : i] = pool[task.coeffs + i];
        __syncwarp();
        inverseTransform(M, sA[warp], sB[warp], log2n, task.trType != 0, bitDepth, lane);
        __syncwarp();
        int sd, sp;
        Sample *dst = hvbBlockPtrW<Sample>(planes, task.dst, sd);
        const Sample *pred = hvbBlockPtr<Sample>(planes, task.pred, sp);
        for (int i = lane; i < count; i += 32)
        {
            const int y = i >> log2n, x = i & (nn - 1);
            dst[y * sd + x] = (Sample)hvbClip3(0, maxv, (int)pred[y * sp + x] + sA[warp][i]);
        }
        __syncwarp();
    }
}

// ---- the fused TU pipeline, in three stages -----------------------------------------------------
//
//   front  (warp per TU)    residual, SSD of the prediction, forward transform; then either the plain
//                           quantiser (levels + cbf, done) or the cooperative RDOQ pre-pass (coefficients
//                           parked in a scratch pool, zero levels written);
//   rdoq   (thread per TU)  the level-coding recurrence (hvb_rdoq.cuh) -- a warp advances 32 TUs;
//   back   (warp per TU)    dequantise, inverse transform, add, clip, SSD of the reconstruction.
//
// The split exists because RDOQ is serial per block: with a warp per block 31 lanes idled through it and the
// ncu capture showed it at 10x the cost of everything else in the pipeline.

// The serial stage gives a thread to each block, so a warp runs as long as its longest block and diverges wherever
// the 32 walks differ.  The work of a walk grows with the position of the first non-zero level (lastSp), which the
// front stage knows: it files every RDOQ block under a bucket of similar lastSp (two per octave), a small kernel
// turns the bucket counts into an ordering (longest first), and the serial stage takes its blocks in that order.
constexpr int kRdoqBuckets = 24;
__device__ __forceinline__ int rdoqBucket(int lastSp)
{
    const int v = lastSp + 1, e = 31 - __clz(v);
    return 2 * e + (e ? (v >> (e - 1)) & 1 : 0);
}

// Compaction of the RDOQ workspace: exclusive prefix sum of the element counts of a batch's RDOQ blocks.
// compact[t] is the offset inside t's 1024-task chunk, chunkBase[t >> 10] the chunk's base (after scanBlocksKernel).
struct TuCount
{
    const hvb_tu_task *tasks;
    __device__ int operator()(int t) const { return (tasks[t].flags & 1) ? 1 << (2 * tasks[t].log2n) : 0; }
};
struct RdoqCount
{
    const hvb_rdoq_task *tasks;
    __device__ int operator()(int t) const { return 1 << (2 * tasks[t].log2n); }
};

template <class Count>
__global__ void __launch_bounds__(1024) scanLocalKernel(Count count, int n, int *__restrict__ compact, int *__restrict__ chunkBase)
{
    __shared__ int sWarp[32];
    const int t = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = t < n ? count(t) : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        int w = sWarp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        sWarp[lane] = w; // inclusive over the warps
    }
    __syncthreads();
    const int base = warp ? sWarp[warp - 1] : 0;
    if (t < n) compact[t] = base + incl - v;
    if (threadIdx.x == 1023) chunkBase[blockIdx.x] = base + incl;
}

// exclusive scan of the chunk totals, in place (one block; chunks beyond 1024 are walked with a running total)
__global__ void __launch_bounds__(1024) scanBlocksKernel(int *__restrict__ chunkBase, int chunks)
{
    __shared__ int sWarp[32];
    __shared__ int sRun;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sRun = 0;
    __syncthreads();
    for (int first = 0; first < chunks; first += 1024)
    {
        const int i = first + threadIdx.x;
        const int v = i < chunks ? chunkBase[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            int w = sWarp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            sWarp[lane] = w;
        }
        __syncthreads();
        const int run = sRun;
        const int base = warp ? sWarp[warp - 1] : 0;
        if (i < chunks) chunkBase[i] = run + base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) sRun = run + base + incl;
        __syncthreads();
    }
}

// counts[kRdoqBuckets] filled by the front stage; cursors[kRdoqBuckets] zero on entry.  Ranks are taken inside the
// block first (shared-memory atomics), then one global atomic per bucket and block reserves the block's range.
__global__ void __launch_bounds__(256)
    tuOrderKernel(const HvbRdoqMid *__restrict__ mids, int n, const int *__restrict__ counts, int *__restrict__ cursors, int *__restrict__ order)
{
    __shared__ int sOffset[kRdoqBuckets], sLocal[kRdoqBuckets];
    if (threadIdx.x < kRdoqBuckets) sLocal[threadIdx.x] = 0;
    if (threadIdx.x == 0)
    {
        int acc = 0;
        for (int b = kRdoqBuckets - 1; b >= 0; --b)
        {
            sOffset[b] = acc;
            acc += counts[b];
        }
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int key = t < n ? mids[t].reserved : -1;
    const int rank = key >= 0 ? atomicAdd(&sLocal[key], 1) : 0;
    __syncthreads();
    if (threadIdx.x < kRdoqBuckets && sLocal[threadIdx.x])
        sOffset[threadIdx.x] += atomicAdd(&cursors[threadIdx.x], sLocal[threadIdx.x]);
    __syncthreads();
    if (key >= 0) order[sOffset[key] + rank] = t;
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    tuFrontKernel(const HvbPlane *__restrict__ planes, int16_t *__restrict__ pool, int16_t *__restrict__ coefTmp,
                  const hvb_rdoq_ctx *__restrict__ rdoqCtx, const hvb_tu_task *__restrict__ tasks, int n, hvb_tu_result *__restrict__ out,
                  HvbRdoqMid *__restrict__ mids, int *__restrict__ bucketCounts, int bitDepth, const int *__restrict__ compact,
                  const int *__restrict__ chunkBase, int rdoqCtxCount, unsigned poolCount)
{
    __shared__ Matrices M;
    __shared__ __align__(16) int16_t sA[kWarps][kBlk];
    __shared__ __align__(16) int16_t sB[kWarps][kBlk];
    initMatrices(M);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * kWarps;
    for (int t = blockIdx.x * kWarps + warp; t < n; t += warpsTotal)
    {
        const hvb_tu_task task = tasks[t];
        const int log2n = task.log2n, nn = 1 << log2n, count = nn * nn;
        const bool dst = task.trType != 0;
        // a task that points outside the pool or at a context snapshot that was never uploaded is not run:
        // its result carries status = -1 and the later stages skip it
        if (log2n < 2 || log2n > 5 || task.levels < 0 || (unsigned)task.levels + (unsigned)count > poolCount ||
            ((task.flags & 1) && (unsigned)task.rdoq_ctx >= (unsigned)rdoqCtxCount))
        {
            if (lane == 0)
            {
                HvbRdoqMid bad;
                bad.lastSp = -2;
                bad.reserved = -1;
                bad.totalDist0 = bad.tailDist0 = 0;
                mids[t] = bad;
                hvb_tu_result r;
                r.ssd = r.ssdPred = 0;
                r.cbf = 0;
                r.status = -1;
                r.sadQuad[0] = r.sadQuad[1] = r.sadQuad[2] = r.sadQuad[3] = 0;
                out[t] = r;
            }
            continue;
        }
        int ss, sp;
        const Sample *src = hvbBlockPtr<Sample>(planes, task.src, ss);
        const Sample *pred = hvbBlockPtr<Sample>(planes, task.pred, sp);

        // residual (Reconstruct.cpp:1275-1287) and SSD of the prediction (:856)
        // and the per-quadrant sum of absolute differences (:1268-1287).  4x4 blocks have 2x2 quadrants: with
        // count = 16 a lane holds at most one sample; larger blocks give a lane samples of one column half only
        // (x = i & (nn - 1) and 32 is a multiple of nn or the other way round), split by row.
        unsigned ssdPred = 0, sadTop = 0, sadBottom = 0;
        for (int i = lane; i < count; i += 32)
        {
            const int y = i >> log2n, x = i & (nn - 1);
            const int d = (int)src[y * ss + x] - (int)pred[y * sp + x];
            sA[warp][i] = (int16_t)d;
            ssdPred += (unsigned)(d * d);
            if (y >> (log2n - 1)) sadBottom += (unsigned)abs(d);
            else sadTop += (unsigned)abs(d);
        }
        unsigned sadQuad[4];
        {
            // a lane's column half is fixed: lane & (nn - 1) when nn <= 32 (the only sizes there are)
            const bool right = ((lane & (nn - 1)) >> (log2n - 1)) != 0;
            sadQuad[0] = hvbWarpSumU(right ? 0u : sadTop);
            sadQuad[1] = hvbWarpSumU(right ? sadTop : 0u);
            sadQuad[2] = hvbWarpSumU(right ? 0u : sadBottom);
            sadQuad[3] = hvbWarpSumU(right ? sadBottom : 0u);
        }
        __syncwarp();
        forwardTransform(M, sA[warp], sB[warp], log2n, dst, bitDepth, lane); // sA = coefficients
        __syncwarp();
        ssdPred = hvbWarpSumU(ssdPred);
        if (sizeof(Sample) == 2) ssdPred >>= 4;

        HvbRdoqMid mid;
        int cbf = 0;
        if (task.flags & 1)
        {
            int16_t *tmp = coefTmp + chunkBase[t >> 10] + compact[t];
            for (int i = lane; i < count; i += 32) tmp[i] = sA[warp][i];
            mid = hvbRdoqPrepass(pool + task.levels, sA[warp], rdoqCtx + task.rdoq_ctx, task.qscale, task.qshift, task.iqscale, log2n,
                                 task.cIdx, task.scanIdx, bitDepth, lane);
        }
        else
        {
            const int off = task.qoffset << (task.qshift - 16);
            int any = 0;
            for (int i = lane; i < count; i += 32)
            {
                const int q = quantOne(sA[warp][i], task.qscale, task.qshift, off);
                any |= q;
                pool[task.levels + i] = (int16_t)q;
            }
            cbf = __any_sync(0xffffffffu, any != 0);
            mid.lastSp = -2; // not an RDOQ block: the serial stage skips it
            mid.reserved = 0;
            mid.totalDist0 = mid.tailDist0 = 0;
        }
        if (lane == 0)
        {
            mid.reserved = mid.lastSp >= 0 ? rdoqBucket(mid.lastSp) : -1;
            if (mid.lastSp >= 0) atomicAdd(&bucketCounts[mid.reserved], 1);
            mids[t] = mid;
            hvb_tu_result r;
            r.ssd = 0;
            r.ssdPred = ssdPred;
            r.cbf = cbf;
            r.status = 0;
            r.sadQuad[0] = sadQuad[0];
            r.sadQuad[1] = sadQuad[1];
            r.sadQuad[2] = sadQuad[2];
            r.sadQuad[3] = sadQuad[3];
            out[t] = r;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128)
    tuRdoqKernel(int16_t *__restrict__ pool, const int16_t *__restrict__ coefTmp, HvbCoefRec *__restrict__ recs,
                 const hvb_rdoq_ctx *__restrict__ rdoqCtx, const hvb_tu_task *__restrict__ tasks, int n, hvb_tu_result *__restrict__ out,
                 const HvbRdoqMid *__restrict__ mids, const int *__restrict__ bucketCounts, const int *__restrict__ order, int bitDepth,
                 const int2 *__restrict__ rdoqBits, const int *__restrict__ rdoqLast, const int *__restrict__ compact,
                 const int *__restrict__ chunkBase)
{
    // blocks that are plain-quantised, or whose levels all round to zero (cbf already 0), are not in the ordering
    int total = 0;
    for (int b = 0; b < kRdoqBuckets; ++b) total += bucketCounts[b];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const int t = order[idx];
        const HvbRdoqMid mid = mids[t];
        const hvb_tu_task task = tasks[t];
        const int ws = chunkBase[t >> 10] + compact[t];
        const int c = hvbRdoqThread(pool + task.levels, coefTmp + ws, rdoqCtx + task.rdoq_ctx, mid, task.qscale, task.qshift,
                                    task.iqscale, task.log2n, task.cIdx, task.scanIdx, (task.flags & 2) != 0, (task.flags & 4) != 0, bitDepth,
                                    recs + ws, rdoqBits + (size_t)task.rdoq_ctx * sizeof(hvb_rdoq_ctx),
                                    rdoqLast + (size_t)task.rdoq_ctx * hvb_rdoq::kLastTabPerCtx);
        out[t].cbf = c != 0;
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    tuBackKernel(const HvbPlane *__restrict__ planes, const int16_t *__restrict__ pool, const hvb_tu_task *__restrict__ tasks, int n,
                 hvb_tu_result *__restrict__ out, int bitDepth)
{
    __shared__ Matrices M;
    __shared__ __align__(16) int16_t sA[kWarps][kBlk];
    __shared__ __align__(16) int16_t sB[kWarps][kBlk];
    initMatrices(M);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * kWarps;
    const int maxv = (1 << bitDepth) - 1;
    for (int t = blockIdx.x * kWarps + warp; t < n; t += warpsTotal)
    {
        const hvb_tu_task task = tasks[t];
        const int log2n = task.log2n, nn = 1 << log2n, count = nn * nn;
        const bool dst = task.trType != 0;
        const int cbf = out[t].cbf;
        if (out[t].status == -1) continue; // rejected by the front stage
        int ss, sp, sr;
        const Sample *src = hvbBlockPtr<Sample>(planes, task.src, ss);
        const Sample *pred = hvbBlockPtr<Sample>(planes, task.pred, sp);
        Sample *rec = hvbBlockPtrW<Sample>(planes, task.rec, sr);
        if (cbf)
        {
            const bool tr = inverseWantsTransposed(log2n) && !dst;
            for (int i = lane; i < count; i += 32)
            {
                const int v = dequantOne(pool[task.levels + i], task.iqscale, task.iqshift); // Reconstruct.cpp:822-826
                sA[warp][tr ? ((i & (nn - 1)) << log2n) + (i >> log2n) : i] = (int16_t)v;
            }
            __syncwarp();
            inverseTransform(M, sA[warp], sB[warp], log2n, dst, bitDepth, lane);
            __syncwarp();
        }
        // all-zero levels: the inverse transform of a zero block is zero, the reconstruction is the prediction
        unsigned ssd = 0;
        for (int i = lane; i < count; i += 32)
        {
            const int y = i >> log2n, x = i & (nn - 1);
            const int r = cbf ? hvbClip3(0, maxv, (int)pred[y * sp + x] + sA[warp][i]) : (int)pred[y * sp + x];
            rec[y * sr + x] = (Sample)r;
            const int d = (int)src[y * ss + x] - r;
            ssd += (unsigned)(d * d);
        }
        ssd = hvbWarpSumU(ssd);
        if (sizeof(Sample) == 2) ssd >>= 4;
        if (lane == 0) out[t].ssd = ssd;
        __syncwarp();
    }
}

// ---- the latency form of the chain: one launch, a warp per block ---------------------------------------------------------
// The three-stage form above is built for whole frames: blocks are sorted by the length of their serial walk, 32 walks
// share a warp, the stages hand over through global scratch.  A caller that waits for a handful of blocks (the batched
// encoder: a CU's transform tree, one intra candidate) pays seven launches and the device's latency on every global
// record for that.  Here a block stays with ONE warp from the prediction to the reconstruction: source and prediction
// tiles are fetched once, with every row requested before the first is used (the prediction may sit in page-locked host
// memory: one trip across the bus instead of one per row), the coefficients, the levels and the walk's 32-byte records
// live in shared memory, lane 0 runs the same hvbRdoqThread, and the back stage continues from the levels in place.
// Same device functions, same arithmetic: results are identical to the staged form's (tests run both).
template <typename Sample>
__device__ __forceinline__ void loadTile(Sample *tile, const Sample *g, int stride, int log2n, int lane)
{
    const int nn = 1 << log2n, rowBytes = nn * (int)sizeof(Sample);
    if (((reinterpret_cast<uintptr_t>(g) | (uintptr_t)(stride * (int)sizeof(Sample))) & 3) == 0)
    {
        // words: every load of a batch of eight is requested before the first is stored
        const int wpr = rowBytes >> 2, words = wpr << log2n;
        uint32_t *dst = reinterpret_cast<uint32_t *>(tile);
        const char *base = reinterpret_cast<const char *>(g);
        const intptr_t pitch = (intptr_t)stride * (int)sizeof(Sample);
        for (int first = 0; first < words; first += 256)
        {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int i = first + j * 32 + lane;
                if (i < words) v[j] = *reinterpret_cast<const uint32_t *>(base + (intptr_t)(i / wpr) * pitch + (i % wpr) * 4);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int i = first + j * 32 + lane;
                if (i < words) dst[i] = v[j];
            }
        }
        return;
    }
    const int count = nn * nn;
    for (int first = 0; first < count; first += 256)
    {
        Sample v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int i = first + j * 32 + lane;
            if (i < count) v[j] = g[(i >> log2n) * stride + (i & (nn - 1))];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int i = first + j * 32 + lane;
            if (i < count) tile[i] = v[j];
        }
    }
}

template <typename Sample>
__global__ void __launch_bounds__(32)
    tuFusedKernel(const HvbPlane *__restrict__ planes, int16_t *__restrict__ pool, const hvb_rdoq_ctx *__restrict__ rdoqCtx,
                  const hvb_tu_task *__restrict__ tasks, int n, hvb_tu_result *__restrict__ out, int bitDepth, int rdoqCtxCount, unsigned poolCount)
{
    __shared__ Matrices M;
    __shared__ __align__(16) int16_t sA[kBlk];
    __shared__ __align__(16) int16_t sB[kBlk];
    __shared__ __align__(16) int16_t sLev[kBlk];
    __shared__ __align__(16) Sample sSrc[kBlk];
    __shared__ __align__(16) Sample sPred[kBlk];
    __shared__ __align__(16) HvbCoefRec sRec[kBlk];
    // the block's context snapshot and what the walk reads of it, derived here (rdoqBitsKernel / rdoqLastKernel do the same for
    // the staged form when snapshots are uploaded): the snapshots may sit in page-locked host memory (hvb_rdoq_contexts_wrap)
    __shared__ __align__(16) hvb_rdoq_ctx sCtx;
    __shared__ int2 sBits[sizeof(hvb_rdoq_ctx)];
    __shared__ int sLast[hvb_rdoq::kLastTabPerCtx];
    initMatrices(M);
    __syncthreads();
    const int lane = threadIdx.x;
    const int maxv = (1 << bitDepth) - 1;
    for (int t = blockIdx.x; t < n; t += gridDim.x)
    {
        const hvb_tu_task task = tasks[t];
        const int log2n = task.log2n, nn = 1 << log2n, count = nn * nn;
        const bool dst = task.trType != 0;
        hvb_tu_result r;
        r.ssd = r.ssdPred = 0;
        r.cbf = 0;
        r.status = 0;
        r.sadQuad[0] = r.sadQuad[1] = r.sadQuad[2] = r.sadQuad[3] = 0;
        if (log2n < 2 || log2n > 5 || task.levels < 0 || (unsigned)task.levels + (unsigned)count > poolCount ||
            ((task.flags & 1) && (unsigned)task.rdoq_ctx >= (unsigned)rdoqCtxCount))
        {
            r.status = -1; // as the staged form: rejected, nothing computed
            if (lane == 0) out[t] = r;
            continue;
        }
        int ss, sp, sr;
        const Sample *src = hvbBlockPtr<Sample>(planes, task.src, ss);
        const Sample *pred = hvbBlockPtr<Sample>(planes, task.pred, sp);
        Sample *rec = hvbBlockPtrW<Sample>(planes, task.rec, sr);
        loadTile<Sample>(sPred, pred, sp, log2n, lane);
        loadTile<Sample>(sSrc, src, ss, log2n, lane);
        __syncwarp();

        // front: residual, SSD of the prediction, per-quadrant sums of absolute differences (as tuFrontKernel)
        unsigned ssdPred = 0, sadTop = 0, sadBottom = 0;
        for (int i = lane; i < count; i += 32)
        {
            const int y = i >> log2n;
            const int d = (int)sSrc[i] - (int)sPred[i];
            sA[i] = (int16_t)d;
            ssdPred += (unsigned)(d * d);
            if (y >> (log2n - 1)) sadBottom += (unsigned)abs(d);
            else sadTop += (unsigned)abs(d);
        }
        {
            const bool right = ((lane & (nn - 1)) >> (log2n - 1)) != 0;
            r.sadQuad[0] = hvbWarpSumU(right ? 0u : sadTop);
            r.sadQuad[1] = hvbWarpSumU(right ? sadTop : 0u);
            r.sadQuad[2] = hvbWarpSumU(right ? 0u : sadBottom);
            r.sadQuad[3] = hvbWarpSumU(right ? sadBottom : 0u);
        }
        __syncwarp();
        forwardTransform(M, sA, sB, log2n, dst, bitDepth, lane); // sA = coefficients
        __syncwarp();
        ssdPred = hvbWarpSumU(ssdPred);
        if (sizeof(Sample) == 2) ssdPred >>= 4;
        r.ssdPred = ssdPred;

        int cbf = 0;
        if (task.flags & 1)
        {
            {
                static_assert(sizeof(hvb_rdoq_ctx) % 4 == 0, "snapshot copied as words");
                const uint32_t *g = reinterpret_cast<const uint32_t *>(rdoqCtx + task.rdoq_ctx);
                for (int i = lane; i < (int)(sizeof(hvb_rdoq_ctx) / 4); i += 32) reinterpret_cast<uint32_t *>(&sCtx)[i] = g[i];
                __syncwarp();
                for (int i = lane; i < (int)sizeof(hvb_rdoq_ctx); i += 32)
                {
                    const uint8_t state = reinterpret_cast<const uint8_t *>(&sCtx)[i];
                    sBits[i] = make_int2(hvb_rdoq::kEntropyBits[state >> 1], hvb_rdoq::kEntropyBits[(state >> 1) ^ 1]);
                }
                // the 20 last-position prefix rates of this block's plane class and size, where hvbRdoqThread looks for them
                const int chroma = task.cIdx ? 1 : 0;
                if (lane < 20) sLast[(chroma * 4 + (log2n - 2)) * 20 + lane] = hvb_rdoq::lastPrefixRate(sCtx, lane >= 10, lane % 10, chroma, log2n);
                __syncwarp();
            }
            const HvbRdoqMid mid = hvbRdoqPrepass(sLev, sA, &sCtx, task.qscale, task.qshift, task.iqscale, log2n, task.cIdx,
                                                  task.scanIdx, bitDepth, lane);
            __syncwarp();
            int c = 0;
            if (lane == 0 && mid.lastSp >= 0)
                c = hvbRdoqThread(sLev, sA, &sCtx, mid, task.qscale, task.qshift, task.iqscale, log2n, task.cIdx, task.scanIdx,
                                  (task.flags & 2) != 0, (task.flags & 4) != 0, bitDepth, sRec, sBits, sLast);
            __syncwarp();
            cbf = __shfl_sync(0xffffffffu, c, 0) != 0;
        }
        else
        {
            const int off = task.qoffset << (task.qshift - 16);
            int any = 0;
            for (int i = lane; i < count; i += 32)
            {
                const int q = quantOne(sA[i], task.qscale, task.qshift, off);
                any |= q;
                sLev[i] = (int16_t)q;
            }
            cbf = __any_sync(0xffffffffu, any != 0);
        }
        __syncwarp();
        for (int i = lane; i < count; i += 32) pool[task.levels + i] = sLev[i];
        r.cbf = cbf;

        // back: dequantise, inverse transform, add, clip, SSD (as tuBackKernel)
        if (cbf)
        {
            const bool tr = inverseWantsTransposed(log2n) && !dst;
            for (int i = lane; i < count; i += 32)
            {
                const int v = dequantOne(sLev[i], task.iqscale, task.iqshift);
                sA[tr ? ((i & (nn - 1)) << log2n) + (i >> log2n) : i] = (int16_t)v;
            }
            __syncwarp();
            inverseTransform(M, sA, sB, log2n, dst, bitDepth, lane);
            __syncwarp();
        }
        unsigned ssd = 0;
        for (int i = lane; i < count; i += 32)
        {
            const int y = i >> log2n, x = i & (nn - 1);
            const int v = cbf ? hvbClip3(0, maxv, (int)sPred[i] + sA[i]) : (int)sPred[i];
            rec[y * sr + x] = (Sample)v;
            const int d = (int)sSrc[i] - v;
            ssd += (unsigned)(d * d);
        }
        ssd = hvbWarpSumU(ssd);
        if (sizeof(Sample) == 2) ssd >>= 4;
        r.ssd = ssd;
        if (lane == 0) out[t] = r;
        __syncwarp();
    }
}

// Rdoq::runQuantisation alone, on pool coefficients: cooperative pre-pass, then one thread per block
__global__ void __launch_bounds__(kWarps * 32)
    rdoqPrepassKernel(int16_t *__restrict__ pool, const hvb_rdoq_ctx *__restrict__ rdoqCtx, const hvb_rdoq_task *__restrict__ tasks, int n,
                      HvbRdoqMid *__restrict__ mids, int32_t *__restrict__ cbf, int bitDepth)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarps;
    for (int t = blockIdx.x * kWarps + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_rdoq_task task = tasks[t];
        const HvbRdoqMid mid = hvbRdoqPrepass(pool + task.dst, pool + task.src, rdoqCtx + task.rdoq_ctx, task.qscale, task.qshift,
                                              task.iqscale, task.log2n, task.cIdx, task.scanIdx, bitDepth, lane);
        if (lane == 0)
        {
            mids[t] = mid;
            cbf[t] = 0;
        }
    }
}

__global__ void __launch_bounds__(128)
    rdoqThreadKernel(int16_t *__restrict__ pool, HvbCoefRec *__restrict__ recs, const hvb_rdoq_ctx *__restrict__ rdoqCtx,
                     const hvb_rdoq_task *__restrict__ tasks, int n, const HvbRdoqMid *__restrict__ mids, int32_t *__restrict__ cbf,
                     int bitDepth, const int2 *__restrict__ rdoqBits, const int *__restrict__ rdoqLast, const int *__restrict__ compact,
                     const int *__restrict__ chunkBase)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    {
        const HvbRdoqMid mid = mids[t];
        if (mid.lastSp < 0) continue;
        const hvb_rdoq_task task = tasks[t];
        const int c = hvbRdoqThread(pool + task.dst, pool + task.src, rdoqCtx + task.rdoq_ctx, mid, task.qscale, task.qshift, task.iqscale,
                                    task.log2n, task.cIdx, task.scanIdx, (task.flags & 2) != 0, (task.flags & 4) != 0, bitDepth,
                                    recs + chunkBase[t >> 10] + compact[t], rdoqBits + (size_t)task.rdoq_ctx * sizeof(hvb_rdoq_ctx),
                                    rdoqLast + (size_t)task.rdoq_ctx * hvb_rdoq::kLastTabPerCtx);
        cbf[t] = c != 0;
    }
}

// {bits(bin 0), bits(bin 1)} for every state byte of the uploaded context snapshots
__global__ void rdoqBitsKernel(const hvb_rdoq_ctx *__restrict__ snapshots, int2 *__restrict__ bits, int bytes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bytes) return;
    const uint8_t state = reinterpret_cast<const uint8_t *>(snapshots)[i];
    bits[i] = make_int2(hvb_rdoq::kEntropyBits[state >> 1], hvb_rdoq::kEntropyBits[(state >> 1) ^ 1]);
}

// the last-position prefix rates of every uploaded snapshot (hvb_rdoq::lastPrefixRate): [ctx][cIdx != 0][log2 - 2][x 0..9, y 10..19]
__global__ void rdoqLastKernel(const hvb_rdoq_ctx *__restrict__ snapshots, int *__restrict__ table, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count * hvb_rdoq::kLastTabPerCtx) return;
    const int c = i / hvb_rdoq::kLastTabPerCtx, r = i - c * hvb_rdoq::kLastTabPerCtx;
    const int chroma = r / 80, log2 = 2 + (r % 80) / 20, e = r % 20;
    table[i] = hvb_rdoq::lastPrefixRate(snapshots[c], e >= 10, e % 10, chroma, log2);
}

} // namespace
int hvbLaunchRdoqBits(hvb_context *ctx, int first, int count)
{
    const int bytes = count * (int)sizeof(hvb_rdoq_ctx);
    rdoqBitsKernel<<<(bytes + 255) / 256, 256, 0, ctx->stream>>>(ctx->rdoqCtx + first, ctx->rdoqBits + (size_t)first * sizeof(hvb_rdoq_ctx), bytes);
    HVB_LAUNCH_CHECK(ctx, "rdoqBitsKernel");
    const int entries = count * hvb_rdoq::kLastTabPerCtx;
    rdoqLastKernel<<<(entries + 255) / 256, 256, 0, ctx->stream>>>(ctx->rdoqCtx + first, ctx->rdoqLast + (size_t)first * hvb_rdoq::kLastTabPerCtx, count);
    HVB_LAUNCH_CHECK(ctx, "rdoqLastKernel");
    return HVB_OK;
}
namespace {

// scan tables of hvb_rdoq.cuh, once per device
int initRdoqTables(hvb_context *ctx)
{
    static bool done[64] = {};
    if (ctx->device < 64 && done[ctx->device]) return HVB_OK;
    std::vector<short> host(4 * 3 * 1024, 0);
    for (int log2 = 2; log2 <= 5; ++log2)
        for (int scanIdx = 0; scanIdx < 3; ++scanIdx)
            for (int sp = 0; sp < (1 << (2 * log2)); ++sp)
                host[((log2 - 2) * 3 + scanIdx) * 1024 + sp] = (short)hvb_rdoq::scanToRaster(log2, scanIdx, sp);
    cudaError_t e = cudaMemcpyToSymbol(hvb_rdoq::gScanTable, host.data(), host.size() * sizeof(short));
    if (e != cudaSuccess) return hvbCuda(ctx, e, "rdoq scan tables");
    if (ctx->device < 64) done[ctx->device] = true;
    return HVB_OK;
}

// scratch layout of the pipeline: [coefTmp: elems int16][recs: elems HvbCoefRec][mids: n HvbRdoqMid][buckets][order: n]
// [compact: n][chunkBase: n/1024 + 1], elems = the RDOQ blocks' elements of this batch (an upper bound when the tasks
// live on the device: the host cannot add them up without a round trip)
struct ChainScratch
{
    int16_t *coefTmp;
    HvbCoefRec *recs;
    HvbRdoqMid *mids;
    int *buckets; // [2][kRdoqBuckets]: counts, cursors
    int *order;   // [n]
    int *compact; // [n]
    int *chunkBase;
    int chunks;
};

int chainScratch(hvb_context *ctx, size_t elems, size_t n, bool needCoefTmp, ChainScratch *cs)
{
    const size_t coefBytes = needCoefTmp ? ((elems * sizeof(int16_t) + 255) & ~size_t(255)) : 0;
    const size_t recBytes = (elems * sizeof(HvbCoefRec) + 255) & ~size_t(255);
    const size_t midBytes = (n * sizeof(HvbRdoqMid) + 255) & ~size_t(255);
    const size_t chunks = (n + 1023) / 1024;
    int rc = hvbEnsureScratch(ctx, coefBytes + recBytes + midBytes + 256 + (2 * n + chunks + 1) * sizeof(int));
    if (rc) return rc;
    char *base = static_cast<char *>(ctx->scratch);
    cs->coefTmp = reinterpret_cast<int16_t *>(base);
    cs->recs = reinterpret_cast<HvbCoefRec *>(base + coefBytes);
    cs->mids = reinterpret_cast<HvbRdoqMid *>(base + coefBytes + recBytes);
    cs->buckets = reinterpret_cast<int *>(base + coefBytes + recBytes + midBytes);
    cs->order = cs->buckets + 64;
    cs->compact = cs->order + n;
    cs->chunkBase = cs->compact + n;
    cs->chunks = (int)chunks;
    return HVB_OK;
}

// elements the RDOQ workspace must hold for a batch
template <class Task, class F>
size_t workspaceElems(hvb_context *ctx, const Task *tasks, int n, hvb_mem mem, F count)
{
    size_t bound = (size_t)n * 1024;
    if (mem == HVB_HOST)
    {
        bound = 0;
        for (int i = 0; i < n; ++i) bound += count(tasks[i]);
    }
    else if (bound > ctx->coeffPoolCount)
        bound = ctx->coeffPoolCount; // distinct blocks of one pool cannot add up to more
    return bound;
}

int gridWarps(hvb_context *ctx, int n, int warps, int perSm)
{
    const int blocks = (n + warps - 1) / warps;
    const int cap = ctx->smCount * perSm;
    return blocks < cap ? blocks : cap;
}

int transformBatch(hvb_context *ctx, const hvb_transform_task *tasks, int n, hvb_mem mem, int inverse)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks) && ctx->coeffPool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    transformKernel<<<gridWarps(ctx, n, kWarps, 4), kWarps * 32, 0, ctx->stream>>>(
        ctx->coeffPool, static_cast<const hvb_transform_task *>(st.dTasks), n, ctx->bitDepth, inverse);
    HVB_LAUNCH_CHECK(ctx, "transformKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

} // namespace

extern "C" int hvb_transform_fwd_batch(hvb_context *ctx, const hvb_transform_task *tasks, int n, hvb_mem mem)
{
    return transformBatch(ctx, tasks, n, mem, 0);
}

extern "C" int hvb_transform_inv_batch(hvb_context *ctx, const hvb_transform_task *tasks, int n, hvb_mem mem)
{
    return transformBatch(ctx, tasks, n, mem, 1);
}

extern "C" int hvb_quantize_batch(hvb_context *ctx, const hvb_quant_task *tasks, int n, int32_t *cbf, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && cbf)) && ctx->coeffPool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, cbf, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    quantKernel<<<gridWarps(ctx, n, 8, 8), 256, 0, ctx->stream>>>(ctx->coeffPool, static_cast<const hvb_quant_task *>(st.dTasks), n,
                                                                  static_cast<int32_t *>(st.dOut), 0);
    HVB_LAUNCH_CHECK(ctx, "quantKernel");
    return hvbStageOut(ctx, cbf, sizeof(int32_t) * n, mem, st);
}

extern "C" int hvb_quantize_inverse_batch(hvb_context *ctx, const hvb_quant_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks) && ctx->coeffPool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    quantKernel<<<gridWarps(ctx, n, 8, 8), 256, 0, ctx->stream>>>(ctx->coeffPool, static_cast<const hvb_quant_task *>(st.dTasks), n,
                                                                  nullptr, 1);
    HVB_LAUNCH_CHECK(ctx, "quantKernel(inverse)");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_inverse_transform_add_batch(hvb_context *ctx, const hvb_ita_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks) && ctx->coeffPool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_ita_task *>(st.dTasks);
    if (ctx->bps == 1)
        itaKernel<uint8_t><<<gridWarps(ctx, n, kWarps, 4), kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, dT, n, ctx->bitDepth);
    else
        itaKernel<uint16_t><<<gridWarps(ctx, n, kWarps, 4), kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, dT, n, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "itaKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_tu_chain_batch(hvb_context *ctx, const hvb_tu_task *tasks, int n, hvb_tu_result *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)) && ctx->coeffPool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    int rc = initRdoqTables(ctx);
    if (rc) return rc;
    if (n <= ctx->tuFusedMax)
    {
        // the latency form: one launch, no scratch (a batch of RDOQ blocks without uploaded snapshots is rejected block by block)
        HvbStaged st;
        rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(hvb_tu_result) * n, mem, &st);
        if (rc) return rc;
        const auto *dT = static_cast<const hvb_tu_task *>(st.dTasks);
        auto *dO = static_cast<hvb_tu_result *>(st.dOut);
        // levels and snapshots: the context's own device arrays, or the caller's page-locked ones (hvb_coeff_pool_wrap,
        // hvb_rdoq_contexts_wrap: the kernel addresses them across the bus, no copy is enqueued)
        int16_t *levelPool = ctx->coeffWrap ? ctx->coeffWrap : ctx->coeffPool;
        const size_t levelCount = ctx->coeffWrap ? ctx->coeffWrapCount : ctx->coeffPoolCount;
        const unsigned poolCount = (unsigned)(levelCount > 0x7fffffffu ? 0x7fffffffu : levelCount);
        const int grid = n < ctx->smCount * 4 ? n : ctx->smCount * 4;
        const hvb_rdoq_ctx *snapshots = ctx->rdoqWrap ? ctx->rdoqWrap : ctx->rdoqCtx;
        const int ctxCount = ctx->rdoqWrap ? ctx->rdoqWrapCount : (ctx->rdoqCtx ? ctx->rdoqCtxCount : 0);
        if (ctx->bps == 1)
            tuFusedKernel<uint8_t><<<grid, 32, 0, ctx->stream>>>(ctx->dPlanes, levelPool, snapshots, dT, n, dO, ctx->bitDepth, ctxCount, poolCount);
        else
            tuFusedKernel<uint16_t><<<grid, 32, 0, ctx->stream>>>(ctx->dPlanes, levelPool, snapshots, dT, n, dO, ctx->bitDepth, ctxCount, poolCount);
        HVB_LAUNCH_CHECK(ctx, "tuFusedKernel");
        return hvbStageOut(ctx, out, sizeof(hvb_tu_result) * n, mem, st);
    }
    if (ctx->coeffWrap || ctx->rdoqWrap)
        return hvbFail(ctx, HVB_ERR_INVALID, "hvb_tu_chain_batch: wrapped level / snapshot arrays serve the one-launch form only (raise hvb_set_tu_fused_max)");
    ChainScratch cs;
    const size_t elems = workspaceElems(ctx, tasks, n, mem, [](const hvb_tu_task &t) {
        return (t.flags & 1) && t.log2n >= 2 && t.log2n <= 5 ? size_t(1) << (2 * t.log2n) : size_t(0);
    });
    if (mem == HVB_HOST && elems && !ctx->rdoqCtx) return hvbFail(ctx, HVB_ERR_INVALID, "hvb_tu_chain_batch: RDOQ tasks before hvb_rdoq_contexts_upload");
    rc = chainScratch(ctx, elems, (size_t)n, true, &cs);
    if (rc) return rc;
    HvbStaged st;
    rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(hvb_tu_result) * n, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_tu_task *>(st.dTasks);
    auto *dO = static_cast<hvb_tu_result *>(st.dOut);
    const int gridW = gridWarps(ctx, n, kWarps, 8);
    int gridT = (n + 127) / 128;
    if (gridT > ctx->smCount * 16) gridT = ctx->smCount * 16;
    cudaMemsetAsync(cs.buckets, 0, 2 * kRdoqBuckets * sizeof(int), ctx->stream);
    scanLocalKernel<<<cs.chunks, 1024, 0, ctx->stream>>>(TuCount{dT}, n, cs.compact, cs.chunkBase);
    HVB_LAUNCH_CHECK(ctx, "scanLocalKernel");
    scanBlocksKernel<<<1, 1024, 0, ctx->stream>>>(cs.chunkBase, cs.chunks);
    HVB_LAUNCH_CHECK(ctx, "scanBlocksKernel");
    const unsigned poolCount = (unsigned)(ctx->coeffPoolCount > 0x7fffffffu ? 0x7fffffffu : ctx->coeffPoolCount);
    if (ctx->bps == 1)
        tuFrontKernel<uint8_t><<<gridW, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, cs.coefTmp, ctx->rdoqCtx, dT, n, dO, cs.mids,
                                                                       cs.buckets, ctx->bitDepth, cs.compact, cs.chunkBase, ctx->rdoqCtx ? ctx->rdoqCtxCount : 0, poolCount);
    else
        tuFrontKernel<uint16_t><<<gridW, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, cs.coefTmp, ctx->rdoqCtx, dT, n, dO, cs.mids,
                                                                        cs.buckets, ctx->bitDepth, cs.compact, cs.chunkBase, ctx->rdoqCtx ? ctx->rdoqCtxCount : 0, poolCount);
    HVB_LAUNCH_CHECK(ctx, "tuFrontKernel");
    tuOrderKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(cs.mids, n, cs.buckets, cs.buckets + kRdoqBuckets, cs.order);
    HVB_LAUNCH_CHECK(ctx, "tuOrderKernel");
    tuRdoqKernel<<<gridT, 128, 0, ctx->stream>>>(ctx->coeffPool, cs.coefTmp, cs.recs, ctx->rdoqCtx, dT, n, dO, cs.mids, cs.buckets, cs.order,
                                                 ctx->bitDepth, ctx->rdoqBits, ctx->rdoqLast, cs.compact, cs.chunkBase);
    HVB_LAUNCH_CHECK(ctx, "tuRdoqKernel");
    if (ctx->bps == 1)
        tuBackKernel<uint8_t><<<gridW, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, dT, n, dO, ctx->bitDepth);
    else
        tuBackKernel<uint16_t><<<gridW, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, ctx->coeffPool, dT, n, dO, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "tuBackKernel");
    return hvbStageOut(ctx, out, sizeof(hvb_tu_result) * n, mem, st);
}

extern "C" int hvb_rdoq_batch(hvb_context *ctx, const hvb_rdoq_task *tasks, int n, int32_t *cbf, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && cbf)) && ctx->coeffPool && ctx->rdoqCtx);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    int rc = initRdoqTables(ctx);
    if (rc) return rc;
    ChainScratch cs;
    const size_t elems = workspaceElems(ctx, tasks, n, mem, [](const hvb_rdoq_task &t) {
        return t.log2n >= 2 && t.log2n <= 5 ? size_t(1) << (2 * t.log2n) : size_t(0);
    });
    rc = chainScratch(ctx, elems, (size_t)n, false, &cs);
    if (rc) return rc;
    HvbStaged st;
    rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, cbf, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_rdoq_task *>(st.dTasks);
    auto *dC = static_cast<int32_t *>(st.dOut);
    scanLocalKernel<<<cs.chunks, 1024, 0, ctx->stream>>>(RdoqCount{dT}, n, cs.compact, cs.chunkBase);
    HVB_LAUNCH_CHECK(ctx, "scanLocalKernel");
    scanBlocksKernel<<<1, 1024, 0, ctx->stream>>>(cs.chunkBase, cs.chunks);
    HVB_LAUNCH_CHECK(ctx, "scanBlocksKernel");
    int gridT = (n + 127) / 128;
    if (gridT > ctx->smCount * 16) gridT = ctx->smCount * 16;
    rdoqPrepassKernel<<<gridWarps(ctx, n, kWarps, 8), kWarps * 32, 0, ctx->stream>>>(ctx->coeffPool, ctx->rdoqCtx, dT, n, cs.mids, dC, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "rdoqPrepassKernel");
    rdoqThreadKernel<<<gridT, 128, 0, ctx->stream>>>(ctx->coeffPool, cs.recs, ctx->rdoqCtx, dT, n, cs.mids, dC, ctx->bitDepth, ctx->rdoqBits, ctx->rdoqLast, cs.compact, cs.chunkBase);
    HVB_LAUNCH_CHECK(ctx, "rdoqThreadKernel");
    return hvbStageOut(ctx, cbf, sizeof(int32_t) * n, mem, st);
}
