// hvb_loopfilter.cu -- the pixel passes of the in-loop filters (deblocking, SAO application) on device-resident reconstructed pictures
// (the first row "next" after the hot path; it gates reference-picture availability,
// turing/TaskEncodeSubstream.cpp:81,91, so keeping it on the device removes the per-CTU reconstruction round trip).
//
// Reference semantics (bit-exact):
//   LoopFilter::Picture::deblock<edgeType>   turing/LoopFilter.h:739-777  (region walk over 8x8 blocks, 4:2:0 chroma rule)
//   LumaBlockEdge   decisions + filters      turing/LoopFilter.h:229-357  (H.265 8.7.2.5.3, .6, .7)
//   ChromaBlockEdge                          turing/LoopFilter.h:359-423  (H.265 8.7.2.5.5, .8)
//   LoopFilter::Block / Ctu                  turing/LoopFilter.h:50-163
//   per-CTU regions                          turing/TaskDeblock.cpp:104-127
//   filterBlockSao, restoreUnfilteredRegions turing/LoopFilter.h:849-1017 (SAO; sao_filter_edge / _band: turing/sao.cpp:35-92)
//
// Mapping.  Edges of one type never share samples (they lie 8 apart and a filter reads 4 and writes 3 samples each side),
// so an edge SEGMENT -- four lines across one 8x8 block boundary -- is a thread's unit of work, with all of its samples
// in registers between one load and one store.  Vertical edges: consecutive threads take consecutive blocks of a block row
// (8 bytes per line each: whole 256-byte rows per warp); horizontal edges: consecutive threads take consecutive groups of
// four columns (one 32-bit word per row each).  Chroma segments (4:2:0: every second luma block edge, strength 2 only) are
// further jobs of the same launch.  The side information (LoopFilter::Block per 8x8 block, the slice's tc / beta offsets
// per CTU) is uploaded per picture; tasks are regions, exactly the arguments of Picture::deblock.
//
// SAO: a thread per four horizontally adjacent samples of a component.  The reference filters a whole CTU, copies the
// disabled 8x8 blocks back and then undoes runs of samples on the CTU's top row / left column / last column / last row
// according to which neighbouring CTUs are available; here that procedure is evaluated per sample (type off, block
// disabled or sample in an undo run: the deblocked value is kept), so a sample is read and written once.
//
// SAO statistics (the encoder's side: EncSao.h:111-284): a thread block per (block, component) task; every thread takes
// interior samples of the block in a block-stride loop, classifies each once for the four edge classes and its band, and
// adds (original - reconstructed) and 1 to shared-memory accumulators (integer atomics: the sums are order-independent,
// hence bit-exact); 104 sums per task go out.  One read of each picture instead of the reference's five passes.
//
// Status: written after the round's GPU budget was spent.  The oracle is pinned against the reference templates
// (tests/test_oracle_pin_loopfilter.py); these kernels' own source is bit-exact against it under host emulation
// (tests/test_host_emulated_loopfilter.py, and tests/test_emulated_gpu_suite_widening.py with device-like alignment
// checks); tests/test_gpu_zz_loopfilter.py is the device parity test and has not yet run on a GPU.
#include "hvb_internal.cuh"

namespace {

__device__ __constant__ uint8_t kBeta[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15,
                                             16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
__device__ __constant__ uint8_t kTc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1,  1,  1,  1,
                                           2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
__device__ __constant__ uint8_t kQpCMid[13] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37};

__device__ __forceinline__ int chromaQp(int qPi) { return qPi < 30 ? qPi : (qPi > 42 ? qPi - 6 : kQpCMid[qPi - 30]); }

// four consecutive samples at a 4-sample-aligned address
__device__ __forceinline__ void load4(const uint8_t *p, int (&v)[4])
{
    const uint32_t w = *reinterpret_cast<const uint32_t *>(p);
    v[0] = w & 0xff;
    v[1] = (w >> 8) & 0xff;
    v[2] = (w >> 16) & 0xff;
    v[3] = w >> 24;
}
__device__ __forceinline__ void load4(const uint16_t *p, int (&v)[4])
{
    const uint2 w = *reinterpret_cast<const uint2 *>(p);
    v[0] = w.x & 0xffff;
    v[1] = w.x >> 16;
    v[2] = w.y & 0xffff;
    v[3] = w.y >> 16;
}
__device__ __forceinline__ void store4(uint8_t *p, const int (&v)[4])
{
    *reinterpret_cast<uint32_t *>(p) = (uint32_t)v[0] | (uint32_t)v[1] << 8 | (uint32_t)v[2] << 16 | (uint32_t)v[3] << 24;
}
__device__ __forceinline__ void store4(uint16_t *p, const int (&v)[4])
{
    *reinterpret_cast<uint2 *>(p) = make_uint2((uint32_t)v[0] | (uint32_t)v[1] << 16, (uint32_t)v[2] | (uint32_t)v[3] << 16);
}

struct Side
{
    int qp;
    bool enabled;
    __device__ __forceinline__ explicit Side(hvb_deblock_block b) : qp(b.data >> 1), enabled(!(b.data & 1)) {}
};

// decisions and filters of one luma segment; s[line][p3 p2 p1 p0 q0 q1 q2 q3] is updated in place
__device__ __forceinline__ void lumaSegment(int (&s)[4][8], int bS, Side P, Side Q, int tcOffsetDiv2, int betaOffsetDiv2, int bitDepth)
{
    const int qPL = (Q.qp + P.qp + 1) >> 1;
    const int beta = kBeta[hvbClip3(0, 51, qPL + 2 * betaOffsetDiv2)] << (bitDepth - 8);
    const int tC = kTc[hvbClip3(0, 53, qPL + 2 * (bS - 1) + 2 * tcOffsetDiv2)] << (bitDepth - 8);
    const int maxv = (1 << bitDepth) - 1;
    const int dp0 = abs(s[0][1] - 2 * s[0][2] + s[0][3]), dp3 = abs(s[3][1] - 2 * s[3][2] + s[3][3]);
    const int dq0 = abs(s[0][6] - 2 * s[0][5] + s[0][4]), dq3 = abs(s[3][6] - 2 * s[3][5] + s[3][4]);
    if (dp0 + dq0 + dp3 + dq3 >= beta) return;
    const bool strong0 = 2 * (dp0 + dq0) < (beta >> 2) && abs(s[0][0] - s[0][3]) + abs(s[0][4] - s[0][7]) < (beta >> 3) &&
                         abs(s[0][3] - s[0][4]) < ((5 * tC + 1) >> 1);
    const bool strong3 = 2 * (dp3 + dq3) < (beta >> 2) && abs(s[3][0] - s[3][3]) + abs(s[3][4] - s[3][7]) < (beta >> 3) &&
                         abs(s[3][3] - s[3][4]) < ((5 * tC + 1) >> 1);
    const bool strong = strong0 && strong3;
    const int sideThreshold = (beta + (beta >> 1)) >> 3;
    const bool dEp = dp0 + dp3 < sideThreshold, dEq = dq0 + dq3 < sideThreshold;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const int p3 = s[k][0], p2 = s[k][1], p1 = s[k][2], p0 = s[k][3], q0 = s[k][4], q1 = s[k][5], q2 = s[k][6], q3 = s[k][7];
        if (strong)
        {
            if (P.enabled)
            {
                s[k][3] = hvbClip3(p0 - 2 * tC, p0 + 2 * tC, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
                s[k][2] = hvbClip3(p1 - 2 * tC, p1 + 2 * tC, (p2 + p1 + p0 + q0 + 2) >> 2);
                s[k][1] = hvbClip3(p2 - 2 * tC, p2 + 2 * tC, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
            }
            if (Q.enabled)
            {
                s[k][4] = hvbClip3(q0 - 2 * tC, q0 + 2 * tC, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
                s[k][5] = hvbClip3(q1 - 2 * tC, q1 + 2 * tC, (p0 + q0 + q1 + q2 + 2) >> 2);
                s[k][6] = hvbClip3(q2 - 2 * tC, q2 + 2 * tC, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
            }
        }
        else
        {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) < tC * 10)
            {
                delta = hvbClip3(-tC, tC, delta);
                if (P.enabled) s[k][3] = hvbClip3(0, maxv, p0 + delta);
                if (Q.enabled) s[k][4] = hvbClip3(0, maxv, q0 - delta);
                if (dEp && P.enabled) s[k][2] = hvbClip3(0, maxv, p1 + hvbClip3(-(tC >> 1), tC >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
                if (dEq && Q.enabled) s[k][5] = hvbClip3(0, maxv, q1 + hvbClip3(-(tC >> 1), tC >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
            }
        }
    }
}

template <typename Sample>
__global__ void __launch_bounds__(256)
    deblockKernel(const HvbPlane *__restrict__ planes, const HvbLoopInfo *__restrict__ info, const hvb_deblock_task *__restrict__ tasks, int n,
                  int bitDepth)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int ti = 0; ti < n; ++ti)
    {
        const hvb_deblock_task t = tasks[ti];
        const HvbLoopInfo li = info[t.pic];
        const int bx0 = t.xBegin / 8, by0 = t.yBegin / 8, nbx = t.xEnd / 8 - bx0, nby = t.yEnd / 8 - by0;
        if (nbx <= 0 || nby <= 0 || !li.blocks) continue;
        const int edge = t.edgeType;
        const int lumaJobs = nbx * nby * 2, jobs = 2 * lumaJobs;
        for (int job = gtid; job < jobs; job += gthreads)
        {
            const bool chroma = job >= lumaJobs;
            const int j = chroma ? job - lumaJobs : job;
            // luma: (block row, position, block) for vertical edges, (block row, block, position) for horizontal ones;
            // chroma: (block row, component, block)
            const int byi = j / (2 * nbx), r = j - byi * 2 * nbx;
            int bxi, sel;
            if (chroma || edge == 0)
            {
                sel = r / nbx;
                bxi = r - sel * nbx;
            }
            else
            {
                bxi = r >> 1;
                sel = r & 1;
            }
            const int bx = bx0 + bxi, by = by0 + byi;
            const hvb_deblock_block qb = li.blocks[(intptr_t)by * li.blockStride + bx];
            const int bS = (qb.packedBs >> (4 * edge + (chroma ? 0 : 2 * sel))) & 3;
            if (chroma ? (bS != 2 || ((edge == 0 ? bx : by) & 1)) : bS == 0) continue;
            const hvb_deblock_block pb = li.blocks[edge == 0 ? (intptr_t)by * li.blockStride + bx - 1 : (intptr_t)(by - 1) * li.blockStride + bx];
            const Side P(pb), Q(qb);
            const hvb_deblock_ctu ctu = li.ctus[li.widthInCtbs * ((by << 3) >> li.ctbLog2) + ((bx << 3) >> li.ctbLog2)];

            if (!chroma)
            {
                const HvbPlane &pl = planes[t.pic * 3];
                Sample *base = reinterpret_cast<Sample *>(pl.base);
                const intptr_t stride = pl.stride;
                int s[4][8];
                if (edge == 0)
                {
                    Sample *line = base + (intptr_t)(8 * by + 4 * sel) * stride + 8 * bx - 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        int a[4], b[4];
                        load4(line + k * stride, a);
                        load4(line + k * stride + 4, b);
#pragma unroll
                        for (int i = 0; i < 4; ++i) s[k][i] = a[i], s[k][4 + i] = b[i];
                    }
                    lumaSegment(s, bS, P, Q, ctu.tc_offset_div2, ctu.beta_offset_div2, bitDepth);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const int a[4] = {s[k][0], s[k][1], s[k][2], s[k][3]}, b[4] = {s[k][4], s[k][5], s[k][6], s[k][7]};
                        store4(line + k * stride, a);
                        store4(line + k * stride + 4, b);
                    }
                }
                else
                {
                    Sample *col = base + (intptr_t)(8 * by - 4) * stride + 8 * bx + 4 * sel;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        int a[4];
                        load4(col + i * stride, a);
#pragma unroll
                        for (int k = 0; k < 4; ++k) s[k][i] = a[k];
                    }
                    lumaSegment(s, bS, P, Q, ctu.tc_offset_div2, ctu.beta_offset_div2, bitDepth);
#pragma unroll
                    for (int i = 1; i < 7; ++i)
                    {
                        const int a[4] = {s[0][i], s[1][i], s[2][i], s[3][i]};
                        store4(col + i * stride, a);
                    }
                }
            }
            else
            {
                // one segment of four chroma lines per 8-sample luma edge; c = 1 + sel
                const HvbPlane &pl = planes[t.pic * 3 + 1 + sel];
                Sample *base = reinterpret_cast<Sample *>(pl.base);
                const intptr_t stride = pl.stride;
                const int QpC = chromaQp(((Q.qp + P.qp + 1) >> 1) + (sel ? t.crQpOffset : t.cbQpOffset));
                const int tC = kTc[hvbClip3(0, 53, QpC + 2 + 2 * ctu.tc_offset_div2)] << (bitDepth - 8);
                const int maxv = (1 << bitDepth) - 1;
                int c[4][4]; // [line][p1 p0 q0 q1]
                if (edge == 0)
                {
                    Sample *line = base + (intptr_t)(4 * by) * stride + 4 * bx - 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        int a[4], b[4];
                        load4(line + k * stride, a);
                        load4(line + k * stride + 4, b);
                        const int delta = hvbClip3(-tC, tC, (((b[0] - a[3]) << 2) + a[2] - b[1] + 4) >> 3);
                        if (P.enabled) a[3] = hvbClip3(0, maxv, a[3] + delta);
                        if (Q.enabled) b[0] = hvbClip3(0, maxv, b[0] - delta);
                        store4(line + k * stride, a);
                        store4(line + k * stride + 4, b);
                    }
                }
                else
                {
                    Sample *col = base + (intptr_t)(4 * by - 2) * stride + 4 * bx;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        int a[4];
                        load4(col + i * stride, a);
#pragma unroll
                        for (int k = 0; k < 4; ++k) c[k][i] = a[k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const int delta = hvbClip3(-tC, tC, (((c[k][2] - c[k][1]) << 2) + c[k][0] - c[k][3] + 4) >> 3);
                        const int p0 = c[k][1], q0 = c[k][2];
                        if (P.enabled) c[k][1] = hvbClip3(0, maxv, p0 + delta);
                        if (Q.enabled) c[k][2] = hvbClip3(0, maxv, q0 - delta);
                    }
#pragma unroll
                    for (int i = 1; i < 3; ++i)
                    {
                        const int a[4] = {c[0][i], c[1][i], c[2][i], c[3][i]};
                        store4(col + i * stride, a);
                    }
                }
            }
        }
    }
}

// the undo runs of an edge-offset CTU (turing/LoopFilter.h:917-992): counts from the start of the top row / left column and
// from the end of the last column / last row, and where those last ones lie
struct SaoUndo
{
    int T, L, R, B, right, bottom;
    __device__ __forceinline__ SaoUndo(const hvb_sao_ctu &ctu, int eo, int sh, int x0, int y0, int n) : T(0), L(0), R(0), B(0)
    {
        right = ctu.right >> sh;
        bottom = ctu.bottom >> sh;
        const bool availL = (ctu.left >> sh) < x0, availR = right > x0 + n, availT = (ctu.top >> sh) < y0, availB = bottom > y0 + n;
        if (eo == 2)
        {
            if (!ctu.topLeft) ++T, ++L;
            if (!ctu.bottomRight) ++R, ++B;
        }
        if (eo != 1)
        {
            if (!availL) L = n;
            if (!availR) R = n;
        }
        if (eo != 0)
        {
            if (!availT) T = n;
            if (!availB) B = n;
        }
        if (eo == 3)
        {
            if (ctu.topRight) --T, --R;
            if (ctu.bottomLeft) --L, --B;
        }
        right = min(right, x0 + n);
        bottom = min(bottom, y0 + n);
    }
    __device__ __forceinline__ bool undone(int x, int y, int lx, int ly, int n) const
    {
        return (ly == 0 && lx < T) || (lx == 0 && ly < L) || (x == right - 1 && ly >= n - R) || (y == bottom - 1 && lx >= n - B);
    }
};

template <typename Sample>
__global__ void __launch_bounds__(256)
    saoKernel(const HvbPlane *__restrict__ planes, const HvbLoopInfo *__restrict__ info, const hvb_sao_task *__restrict__ tasks, int nTasks,
              int bitDepth)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    const int maxv = (1 << bitDepth) - 1;
    for (int ti = 0; ti < nTasks; ++ti)
    {
        const hvb_sao_task t = tasks[ti];
        const HvbLoopInfo li = info[t.dst_pic];
        const int first = max((int)t.ctuBegin, 0), ctus = min((int)t.ctuEnd, li.saoCount) - first;
        if (ctus <= 0 || !li.sao || !li.blocks) continue;
        for (int c = 0; c < 3; ++c)
        {
            if (!(c ? t.chromaFlag : t.lumaFlag)) continue;
            const int sh = c ? 1 : 0, n = (1 << li.ctbLog2) >> sh, quads = n >> 2, perCtu = n * quads;
            const HvbPlane &sp = planes[t.src_pic * 3 + c], &dp = planes[t.dst_pic * 3 + c];
            const Sample *src = reinterpret_cast<const Sample *>(sp.base);
            Sample *dst = reinterpret_cast<Sample *>(dp.base);
            const int w = dp.width, h = dp.height; // of this component
            for (int job = gtid; job < ctus * perCtu; job += gthreads)
            {
                const int ci = job / perCtu, r = job - ci * perCtu, ly = r / quads, lx = (r - ly * quads) << 2;
                const int addr = first + ci, ry = addr / li.widthInCtbs, rx = addr - ry * li.widthInCtbs;
                const int x0 = rx * n, y0 = ry * n, x = x0 + lx, y = y0 + ly;
                if (x >= w || y >= h) continue;
                const hvb_sao_ctu &ctu = li.sao[addr];
                const int type = ctu.plane[c].typeIdx, eo = ctu.plane[c].classOrBand;
                int v[4];
                load4(src + (intptr_t)y * sp.stride + x, v);
                const bool enabled = !(li.blocks[(intptr_t)((y << sh) >> 3) * li.blockStride + ((x << sh) >> 3)].data & 1);
                if (type == 1 && enabled)
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const int k = ((v[j] >> (bitDepth - 5)) - eo) & 31; // band index relative to sao_band_position
                        if (k < 4) v[j] = hvbClip3(0, maxv, v[j] + ctu.plane[c].offset[k]);
                    }
                }
                else if (type == 2 && enabled)
                {
                    const SaoUndo undo(ctu, eo, sh, x0, y0, n);
                    const int hOff = eo == 1 ? 0 : (eo == 3 ? 1 : -1), vOff = eo == 0 ? 0 : -1;
                    const intptr_t step = (intptr_t)vOff * sp.stride + hOff;
                    const Sample *at = src + (intptr_t)y * sp.stride + x;
                    int out[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        out[j] = v[j];
                        if (x + j < w && !undo.undone(x + j, y, lx + j, ly, n))
                        {
                            const int a = at[j + step], b = at[j - step];
                            const int idx = 2 + (v[j] > a) - (v[j] < a) + (v[j] > b) - (v[j] < b);
                            // edgeIdx 0,1,2 -> 1,2,0 (H.265 8.7.3.2); category 0 (flat) gets no offset
                            const int category = idx == 2 ? 0 : (idx < 2 ? idx + 1 : idx);
                            if (category) out[j] = hvbClip3(0, maxv, v[j] + ctu.plane[c].offset[category - 1]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = out[j];
                }
                store4(dst + (intptr_t)y * dp.stride + x, v);
            }
        }
    }
}

// ---- SAO statistics: one thread block per task -------------------------------------------------------------------------
template <typename Sample>
__global__ void __launch_bounds__(256)
    saoStatsKernel(const HvbPlane *__restrict__ planes, const hvb_sao_stats_task *__restrict__ tasks, int n, hvb_sao_stats *__restrict__ out,
                   int bitDepth)
{
    __shared__ int acc[104]; // hvb_sao_stats as a flat array: edgeE[4][5], edgeCount[4][5], bandE[32], bandCount[32]
    for (int ti = blockIdx.x; ti < n; ti += gridDim.x)
    {
        const hvb_sao_stats_task t = tasks[ti];
        for (int i = threadIdx.x; i < 104; i += blockDim.x) acc[i] = 0;
        __syncthreads();
        const HvbPlane &op = planes[t.org_pic * 3 + t.cIdx], &rp = planes[t.rec_pic * 3 + t.cIdx];
        const Sample *org = reinterpret_cast<const Sample *>(op.base) + (intptr_t)t.y0 * op.stride + t.x0;
        const Sample *rec = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)t.y0 * rp.stride + t.x0;
        const int iw = t.w - 2, ih = t.h - 2; // the interior: rows and columns 1 .. size - 2
        for (int i = threadIdx.x; i < iw * ih; i += blockDim.x)
        {
            const int y = 1 + i / iw, x = 1 + i - (y - 1) * iw;
            const Sample *at = rec + (intptr_t)y * rp.stride + x;
            const int v = at[0], diff = (int)org[(intptr_t)y * op.stride + x] - v;
            const intptr_t up = -(intptr_t)rp.stride;
            // neighbour pairs of the classes: horizontal, vertical, the \ diagonal, the / diagonal (turing/sao.cpp:64-74)
            const int a[4] = {at[-1], at[up], at[up - 1], at[up + 1]}, b[4] = {at[1], at[-up], at[-up + 1], at[-up - 1]};
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
                const int idx = 2 + (v > a[c]) - (v < a[c]) + (v > b[c]) - (v < b[c]);
                const int category = idx == 2 ? 0 : (idx < 2 ? idx + 1 : idx);
                atomicAdd(&acc[c * 5 + category], diff);
                atomicAdd(&acc[20 + c * 5 + category], 1);
            }
            if (x == 1)
            {
                // the reference's class-0 loop visits column 1 a second time, as category 0 (EncSao.h:167-181)
                atomicAdd(&acc[0], diff);
                atomicAdd(&acc[20], 1);
            }
            const int band = v >> (bitDepth - 5);
            atomicAdd(&acc[40 + band], diff);
            atomicAdd(&acc[72 + band], 1);
        }
        __syncthreads();
        int *o = reinterpret_cast<int *>(out + ti);
        for (int i = threadIdx.x; i < 104; i += blockDim.x) o[i] = acc[i];
        __syncthreads();
    }
}

} // namespace

extern "C" int hvb_deblock_info_upload(hvb_context *ctx, int pic, const hvb_deblock_block *blocks, int blockStride, int blockRows,
                                       const hvb_deblock_ctu *ctus, int picWidthInCtbs, int picHeightInCtbs, int ctbLog2)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live);
    HVB_CHECK_ARGS(ctx, blocks && ctus && blockStride > 0 && blockRows > 0 && picWidthInCtbs > 0 && picHeightInCtbs > 0 && ctbLog2 >= 4 && ctbLog2 <= 6);
    cudaSetDevice(ctx->device);
    HvbPicture &p = ctx->pictures[pic];
    const size_t blockBytes = sizeof(hvb_deblock_block) * (size_t)blockStride * blockRows;
    const size_t ctuBytes = sizeof(hvb_deblock_ctu) * (size_t)picWidthInCtbs * picHeightInCtbs;
    cudaError_t e = cudaSuccess;
    if (!ctx->dLoopInfo)
    {
        e = cudaMalloc(&ctx->dLoopInfo, sizeof(HvbLoopInfo) * HVB_MAX_PICTURES);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->dLoopInfo, 0, sizeof(HvbLoopInfo) * HVB_MAX_PICTURES, ctx->stream);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "loop-filter table");
    }
    if (p.lfBytes < blockBytes + ctuBytes + 256)
    {
        cudaStreamSynchronize(ctx->stream);
        if (p.lfInfo) cudaFree(p.lfInfo);
        p.lfInfo = nullptr;
        p.lfBytes = 0;
        e = cudaMalloc(&p.lfInfo, blockBytes + ctuBytes + 256);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "loop-filter side information");
        p.lfBytes = blockBytes + ctuBytes + 256;
    }
    char *dBlocks = static_cast<char *>(p.lfInfo), *dCtus = dBlocks + ((blockBytes + 255) & ~size_t(255));
    int rc = hvbUpload(ctx, dBlocks, blockBytes, blocks, blockBytes, blockBytes, 1, "deblock blocks");
    if (rc) return rc;
    rc = hvbUpload(ctx, dCtus, ctuBytes, ctus, ctuBytes, ctuBytes, 1, "deblock ctus");
    if (rc) return rc;
    HvbLoopInfo &li = ctx->loopInfoHost[pic]; // the host mirror outlives the call; the SAO fields stay as they are
    li.blocks = reinterpret_cast<const hvb_deblock_block *>(dBlocks);
    li.ctus = reinterpret_cast<const hvb_deblock_ctu *>(dCtus);
    li.blockStride = blockStride;
    li.blockRows = blockRows;
    li.widthInCtbs = picWidthInCtbs;
    li.ctbLog2 = ctbLog2;
    e = cudaMemcpyAsync(ctx->dLoopInfo + pic, &li, sizeof(li), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "loop-filter table entry");
    return HVB_OK;
}

extern "C" int hvb_sao_info_upload(hvb_context *ctx, int pic, const hvb_sao_ctu *ctus, int n)
{
    HVB_CHECK_ARGS(ctx, pic >= 0 && pic < HVB_MAX_PICTURES && ctx->pictures[pic].live && ctus && n > 0);
    if (!ctx->dLoopInfo || !ctx->loopInfoHost[pic].blocks)
        return hvbFail(ctx, HVB_ERR_INVALID, "hvb_sao_info_upload before hvb_deblock_info_upload (block records and CTB geometry)");
    cudaSetDevice(ctx->device);
    HvbPicture &p = ctx->pictures[pic];
    const size_t bytes = sizeof(hvb_sao_ctu) * (size_t)n;
    if (p.saoBytes < bytes)
    {
        cudaStreamSynchronize(ctx->stream);
        if (p.saoInfo) cudaFree(p.saoInfo);
        p.saoInfo = nullptr;
        p.saoBytes = 0;
        cudaError_t e = cudaMalloc(&p.saoInfo, bytes);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "SAO records");
        p.saoBytes = bytes;
    }
    int rc = hvbUpload(ctx, p.saoInfo, bytes, ctus, bytes, bytes, 1, "SAO records");
    if (rc) return rc;
    HvbLoopInfo &li = ctx->loopInfoHost[pic];
    li.sao = static_cast<const hvb_sao_ctu *>(p.saoInfo);
    li.saoCount = n;
    cudaError_t e = cudaMemcpyAsync(ctx->dLoopInfo + pic, &li, sizeof(li), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return hvbCuda(ctx, e, "loop-filter table entry");
    return HVB_OK;
}

extern "C" int hvb_sao_batch(hvb_context *ctx, const hvb_sao_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks));
    if (!n) return HVB_OK;
    if (!ctx->dLoopInfo) return hvbFail(ctx, HVB_ERR_INVALID, "hvb_sao_batch before hvb_sao_info_upload");
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_sao_task *>(st.dTasks);
    const int blocks = ctx->smCount * 8;
    if (ctx->bps == 1)
        saoKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, ctx->dLoopInfo, dT, n, ctx->bitDepth);
    else
        saoKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, ctx->dLoopInfo, dT, n, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "saoKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_deblock_batch(hvb_context *ctx, const hvb_deblock_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks));
    if (!n) return HVB_OK;
    if (!ctx->dLoopInfo) return hvbFail(ctx, HVB_ERR_INVALID, "hvb_deblock_batch before hvb_deblock_info_upload");
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_deblock_task *>(st.dTasks);
    const int blocks = ctx->smCount * 8;
    if (ctx->bps == 1)
        deblockKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, ctx->dLoopInfo, dT, n, ctx->bitDepth);
    else
        deblockKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, ctx->dLoopInfo, dT, n, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "deblockKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_sao_stats_batch(hvb_context *ctx, const hvb_sao_stats_task *tasks, int n, hvb_sao_stats *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    static_assert(sizeof(hvb_sao_stats) == 104 * sizeof(int32_t), "saoStatsKernel writes hvb_sao_stats as 104 ints");
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(*out) * n, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_sao_stats_task *>(st.dTasks);
    auto *dO = static_cast<hvb_sao_stats *>(st.dOut);
    const int blocks = min(n, ctx->smCount * 8);
    if (ctx->bps == 1)
        saoStatsKernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    else
        saoStatsKernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "saoStatsKernel");
    return hvbStageOut(ctx, out, sizeof(*out) * n, mem, st);
}

extern "C" int hvb_picture_copy(hvb_context *ctx, int dst_pic, int src_pic)
{
    HVB_CHECK_ARGS(ctx, dst_pic >= 0 && dst_pic < HVB_MAX_PICTURES && src_pic >= 0 && src_pic < HVB_MAX_PICTURES && dst_pic != src_pic);
    const HvbPicture &d = ctx->pictures[dst_pic], &s = ctx->pictures[src_pic];
    HVB_CHECK_ARGS(ctx, d.live && s.live && d.width == s.width && d.height == s.height && d.pad == s.pad);
    cudaSetDevice(ctx->device);
    for (int c = 0; c < 3; ++c)
    {
        if (d.allocBytes[c] != s.allocBytes[c]) return hvbFail(ctx, HVB_ERR_INVALID, "hvb_picture_copy: plane allocations differ");
        cudaError_t e = cudaMemcpyAsync(d.alloc[c], s.alloc[c], s.allocBytes[c], cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) return hvbCuda(ctx, e, "hvb_picture_copy");
    }
    return HVB_OK;
}
