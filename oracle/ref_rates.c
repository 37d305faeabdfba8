/*
 * oracle/ref_rates.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Streaming rates of the UNMODIFIED reference primitives on one host core, for BASELINE.md section 3's per-primitive table:
 * the havoc tables of oracle/_ref/libhavoc_ref.so (C path = the `--asm 0` identity path, and the best JIT instruction set = the
 * `--asm 1` speed path) called through oracle/ref_shim.cpp on disjoint blocks of two planes that are far larger than the last-
 * level cache, i.e. the same access pattern tools/stream_metrics.py gives the B200 kernels.  The reference's own harness
 * (havoc/havoc.cpp:161-211, havoc_test.c:76-91) times in-cache calls in cycles; this reports what a search over whole pictures
 * sees.  Output: one JSON object per line, bytes by SURVEY.md section 8(d)'s formulas.
 *
 * usage: ref_rates [side] [seconds per row]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

void *ref_create(int use_asm);
void ref_destroy(void *h);
unsigned ref_isa(void *h);
int ref_sad(void *h, const void *src, intptr_t ss, const void *ref, intptr_t sr, int w, int hgt, int bps);
void ref_sad_multiref4(void *h, const void *src, intptr_t ss, const void *const ref[4], intptr_t sr, int sad[4], int w, int hgt, int bps);
uint32_t ref_ssd(void *h, const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps);
int ref_hadamard_satd(void *h, const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps);

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv)
{
    const int side = argc > 1 ? atoi(argv[1]) : 8192;
    const double budget = argc > 2 ? atof(argv[2]) : 0.5;
    long long sink = 0;
    for (int bps = 1; bps <= 2; ++bps)
    {
        const size_t bytes = (size_t)side * side * bps;
        uint8_t *a = aligned_alloc(64, bytes + 4096), *b = aligned_alloc(64, bytes + 4096);
        if (!a || !b) return 1;
        unsigned seed = 12345u;
        for (size_t i = 0; i < bytes + 4096; ++i)
        {
            seed = seed * 1664525u + 1013904223u;
            a[i] = (uint8_t)(seed >> 24);
            b[i] = (uint8_t)(seed >> 16);
            if (bps == 2 && (i & 1)) a[i] &= 3, b[i] &= 3; /* 10-bit samples */
        }
        for (int use_asm = 0; use_asm <= 1; ++use_asm)
        {
            void *h = ref_create(use_asm);
            for (int n = 64; n >= 8; n >>= 1)
                for (int kind = 0; kind < 4; ++kind) /* sad, sad4, ssd, satd (8x8 tiles, as measureSatd walks them) */
                    for (int unaligned = 0; unaligned <= 1; ++unaligned)
                    {
                        if (kind == 2 && unaligned) continue; /* SSD compares co-located blocks only (Reconstruct.cpp:351-353) */
                        const int pitch = unaligned ? n + 16 : n, cols = (side - 32) / pitch, rows = side / n, rows4 = rows / 4;
                        const double t0 = now();
                        double t1 = t0;
                        long long calls = 0;
                        int log2n = 0;
                        while ((1 << log2n) < n) ++log2n;
                        for (int pass = 0; t1 - t0 < budget; ++pass)
                        {
                            for (int by = 0; by < (kind == 1 ? rows4 : rows); ++by)
                                for (int bx = 0; bx < cols; ++bx)
                                {
                                    const int dx = unaligned ? 1 + ((bx * 7 + by * 3) % 15) : 0;
                                    const uint8_t *pa = a + ((size_t)by * n * side + (size_t)bx * pitch) * bps;
                                    const uint8_t *pb = b + ((size_t)by * n * side + (size_t)bx * pitch + dx) * bps;
                                    if (kind == 0)
                                        sink += ref_sad(h, pa, side, pb, side, n, n, bps);
                                    else if (kind == 1)
                                    {
                                        const void *refs[4];
                                        int sad[4];
                                        for (int k = 0; k < 4; ++k) refs[k] = pb + (size_t)k * (side / 4) * side * bps;
                                        ref_sad_multiref4(h, pa, side, refs, side, sad, n, n, bps);
                                        sink += sad[0] + sad[3];
                                    }
                                    else if (kind == 2)
                                        sink += ref_ssd(h, pa, side, pb, side, log2n, bps);
                                    else
                                        for (int ty = 0; ty < n; ty += 8)
                                            for (int tx = 0; tx < n; tx += 8)
                                                sink += ref_hadamard_satd(h, pa + ((size_t)ty * side + tx) * bps, side, pb + ((size_t)ty * side + tx) * bps, side, 3, bps);
                                    ++calls;
                                }
                            t1 = now();
                        }
                        const char *names[4] = {"sad", "sad4", "ssd", "satd"};
                        const double per = (kind == 1 ? 5.0 : 2.0) * n * n * bps;
                        printf("{\"primitive\": \"%s\", \"block\": %d, \"bytes_per_sample\": %d, \"layout\": \"%s\", \"tables\": \"%s\", \"isa_mask\": %u, "
                               "\"calls\": %lld, \"ns_per_call\": %.1f, \"GBps_one_core\": %.3f}\n",
                               names[kind], n, bps, unaligned ? "unaligned" : "colocated", use_asm ? "jit" : "c", ref_isa(h), calls,
                               1e9 * (t1 - t0) / calls, per * calls / (t1 - t0) / 1e9);
                        fflush(stdout);
                    }
            ref_destroy(h);
        }
        free(a);
        free(b);
    }
    return (int)(sink & 1) * 0;
}
