// hvb_codeddata.cu -- quantised levels -> the encoder's coded-data residual records, on the device.
//
// Reference semantics (bit-exact):
//   CodedData::storeResidual        turing/CodedData.h:457-517 (SubBlock / Residual word layout :117-270)
//   scans                           turing/ScanOrder.h:32-101, :152-187
//
// After the TU chain the levels of a transform block sit in the coefficient pool as a raster n x n int16 block; what the
// host's CABAC writer and rate estimator read is the serialised form: coded-sub-block flags, then per significant 4x4
// sub-block (last in scan order first) a significance, a greater-than-1 and a sign mask and the magnitudes above 1.  A
// thread serialises one block: a first walk sizes the record, an atomic bump of a cursor reserves room in the record
// region of the pool, a second walk fills it.  Typical blocks shrink from 2 n^2 bytes to a few tens, so the host fetches the
// used part of the region and an (offset, length) pair per block instead of whole blocks.
//
// Status: written after the round's GPU budget was spent; bit-exact under host emulation
// (tests/test_host_emulated_codeddata.py); tests/test_gpu_zz_codeddata.py has not yet run on a GPU.
#include "hvb_internal.cuh"

namespace {

// i-th position of scan `scanIdx` (0 up-right diagonal, 1 horizontal, 2 vertical) of a size x size grid
__device__ __forceinline__ void scanPos(int size, int scanIdx, int i, int &x, int &y)
{
    if (scanIdx == 1)
    {
        x = i % size;
        y = i / size;
        return;
    }
    if (scanIdx == 2)
    {
        x = i / size;
        y = i % size;
        return;
    }
    for (int d = 0;; ++d)
    {
        // anti-diagonal d: the cells x + y = d inside the grid, from the bottom-left end upwards
        const int yTop = d < size ? d : size - 1, count = d < size ? d + 1 : 2 * size - 1 - d;
        if (i < count)
        {
            y = yTop - i;
            x = d - y;
            return;
        }
        i -= count;
    }
}

__global__ void __launch_bounds__(128)
    codedResidualKernel(const int16_t *pool, uint16_t *records, // one allocation: the level blocks and the record region do not overlap
                         const hvb_coded_residual_task *__restrict__ tasks, int n,
                        int recordsBase, int capacityWords, hvb_coded_residual *__restrict__ out, int *__restrict__ cursor)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    for (int ti = gtid; ti < n; ti += gthreads)
    {
        const hvb_coded_residual_task t = tasks[ti];
        const int16_t *levels = pool + t.levels;
        const int size = 1 << t.log2n, grid = size >> 2, subBlocks = grid * grid, header = 1 + (t.log2n == 5 ? 4 : 1);
        int inner[16]; // offsets of the 16 scan positions inside a sub-block
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            int x, y;
            scanPos(4, t.scanIdx, k, x, y);
            inner[k] = y * size + x;
        }
        // first walk: the record's length
        int words = header;
        bool any = false;
        for (int i = 0; i < subBlocks; ++i)
        {
            int sx, sy;
            scanPos(grid, t.scanIdx, i, sx, sy);
            const int16_t *block = levels + (sy * 4) * size + sx * 4;
            int nz = 0, big = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const int v = block[inner[k]];
                nz += v != 0;
                big += v > 1 || v < -1;
            }
            if (nz)
            {
                words += 3 + big;
                any = true;
            }
        }
        if (!any)
        {
            out[ti] = hvb_coded_residual{0, 0}; // cbf = 0: storeResidual writes nothing
            continue;
        }
        const int at = atomicAdd(cursor, words);
        if (at + words > capacityWords)
        {
            out[ti] = hvb_coded_residual{0, -1}; // the region is full; the total reported with out[n] says by how much
            continue;
        }
        // second walk: fill
        uint16_t *rec = records + recordsBase + at;
        for (int i = 0; i < header; ++i) rec[i] = 0;
        uint16_t *p = rec + header;
        for (int i = subBlocks - 1; i >= 0; --i)
        {
            int sx, sy;
            scanPos(grid, t.scanIdx, i, sx, sy);
            const int16_t *block = levels + (sy * 4) * size + sx * 4;
            unsigned sig = 0, greater1 = 0, sign = 0;
            int count = 0;
#pragma unroll
            for (int k = 15; k >= 0; --k)
            {
                const int v = block[inner[k]];
                if (v)
                {
                    sig |= 1u << (15 - k);
                    if (v < 0) sign |= 1u << (15 - k);
                    if (v > 1 || v < -1)
                    {
                        greater1 |= 1u << (15 - k);
                        p[3 + count++] = (uint16_t)abs(v);
                    }
                }
            }
            if (!sig) continue;
            rec[1 + (i >> 4)] |= (uint16_t)(1u << (i & 15));
            p[0] = (uint16_t)sig;
            p[1] = (uint16_t)greater1;
            p[2] = (uint16_t)sign;
            p += 3 + count;
        }
        out[ti] = hvb_coded_residual{recordsBase + at, words};
    }
}

// out[n]: where the used part of the region ends, and whether it overflowed
__global__ void codedResidualTotalKernel(hvb_coded_residual *__restrict__ out, int n, int recordsBase, int capacityWords, const int *__restrict__ cursor)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        const int used = *cursor;
        out[n] = hvb_coded_residual{recordsBase + (used < capacityWords ? used : capacityWords), used > capacityWords ? used - capacityWords : 0};
    }
}

} // namespace

extern "C" int hvb_coded_residual_batch(hvb_context *ctx, const hvb_coded_residual_task *tasks, int n, int32_t recordsBase, int32_t capacityWords,
                                        hvb_coded_residual *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)) && recordsBase >= 0 && capacityWords >= 0);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    int rc = hvbEnsureCoeffPool(ctx, (size_t)recordsBase + capacityWords);
    if (rc) return rc;
    HvbStaged st;
    rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(*out) * (n + 1), mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_coded_residual_task *>(st.dTasks);
    auto *dO = static_cast<hvb_coded_residual *>(st.dOut);
    int *cursor = ctx->workCursors + 3;
    cudaMemsetAsync(cursor, 0, sizeof(int), ctx->stream);
    const int blocks = min((n + 127) / 128, ctx->smCount * 16);
    codedResidualKernel<<<blocks, 128, 0, ctx->stream>>>(ctx->coeffPool, reinterpret_cast<uint16_t *>(ctx->coeffPool), dT, n, recordsBase,
                                                        capacityWords, dO, cursor);
    HVB_LAUNCH_CHECK(ctx, "codedResidualKernel");
    codedResidualTotalKernel<<<1, 32, 0, ctx->stream>>>(dO, n, recordsBase, capacityWords, cursor);
    HVB_LAUNCH_CHECK(ctx, "codedResidualTotalKernel");
    return hvbStageOut(ctx, out, sizeof(*out) * (n + 1), mem, st);
}
