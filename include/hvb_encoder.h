/*
 * hvb_encoder.h -- submission queue between an encoder's worker threads and the batched ABI of hvb.h.
 *
 * The reference encoder runs one task per CTU row and per picture in flight on a thread pool
 * (turing/ThreadPool.cpp:87-103, turing/TaskEncodeSubstream.cpp:150-215): up to PicHeightInCtbs x concurrent-frames
 * threads are inside Search<coding_quadtree>::go at the same time, each issuing one motion search, one PU cost, one
 * intra sweep or one CU's transform blocks at a time.  A session gathers those calls: a worker posts its task and
 * blocks; a dispatcher thread owns the hvb_context, turns everything posted while the previous batch was on the device
 * into one hvb_*_batch call per kind (tasks and results in page-locked memory the kernels address directly, no staging
 * copy), waits for the stream and wakes the workers.  The batch size adapts to the load by itself.
 *
 * Pictures: the session owns a pool of device pictures; a host picture (source or reconstruction) is named by a key
 * (any pointer that identifies it while it lives) and bound to a pool slot by hvbenc_picture().  Reconstructed CTUs go
 * up with hvbenc_upload_rect() as the in-loop filters finish them (turing/TaskSao.cpp:96-153), source pictures once.
 *
 * All entry points may be called concurrently from any thread; each returns 0 or a negative hvb_status after the work
 * has completed (results written, uploads visible to every later submission).
 */
#ifndef HVB_ENCODER_H
#define HVB_ENCODER_H

#include "hvb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hvbenc hvbenc;

/* width x height: luma size of every picture of the session; pool_pictures: device pictures to allocate (least recently
 * used slot is rebound when a new key arrives; an encoder needs its DPB plus two per picture in flight). */
int hvbenc_create(int device, int bytes_per_sample, int bit_depth, int width, int height, int pool_pictures, hvbenc **out);
void hvbenc_destroy(hvbenc *enc);
const char *hvbenc_last_error(hvbenc *enc);

/* Bind `key` to a device picture and return its id (the `pic` of hvb.h's tasks).  fresh != 0: the key names a NEW host
 * picture (its address may have belonged to a dead one): the binding is renewed and nothing of the old content is relied on. */
int hvbenc_picture(hvbenc *enc, const void *key, int fresh, int *pic);
/* copy a rectangle of plane cIdx (x0, y0, w, h in samples of that plane, may extend into the padding) from host memory;
 * `host` points at the rectangle's first sample, stride in samples */
int hvbenc_upload_rect(hvbenc *enc, int pic, int cIdx, const void *host, intptr_t stride, int x0, int y0, int w, int h);
/* several rectangles of one picture in one hand-over (a finished CTU: its luma and both chroma blocks) */
typedef struct
{
    int cIdx;
    const void *host;
    intptr_t stride;
    int x0, y0, w, h;
} hvbenc_rect;
int hvbenc_upload_rects(hvbenc *enc, int pic, const hvbenc_rect *rects, int n);

/* one blocking call each; the session batches across callers */
int hvbenc_me(hvbenc *enc, const hvb_me_task *task, hvb_me_result *out);
int hvbenc_me_bi(hvbenc *enc, const hvb_me_bi_task *task, hvb_me_bi_result *out);
int hvbenc_pu_cost(hvbenc *enc, const hvb_pu_cost_task *tasks, int n, int32_t *out /* [n][3] */);
/* 35-mode sweep of one partition: `neighbours` are the 4n+1 unfiltered reference samples in the reference's order
 * p(-1, n*2-1) .. p(-1,-1) .. p(n*2-1,-1) (index 2n holds p(-1,-1)); the filtered array is derived on the device.
 * task->nb_unfiltered / nb_filtered are ignored. */
int hvbenc_intra_sweep(hvbenc *enc, const hvb_intra_sweep_task *task, const void *neighbours, int32_t *out /* [35] */);

/* The transform blocks of one CU (hvb_tu_chain_batch) with the prediction supplied by the caller and the reconstruction
 * returned to it: block i's prediction is pred[i] (n x n samples, stride pred_stride[i]); on return rec[i] (same geometry,
 * may alias pred[i]) holds the reconstruction, levels[i] the n*n quantised levels (raster), out[i] the SSDs and cbf.
 * tasks[i].src names the source block; tasks[i].pred / .rec / .levels / .rdoq_ctx are filled by the session;
 * `snapshot`: the CABAC context snapshot + lambda all RDOQ blocks of the call use (NULL when none has flags bit 0). */
int hvbenc_tu_chain(hvbenc *enc, hvb_tu_task *tasks, int n, const hvb_rdoq_ctx *snapshot, const void *const *pred, const intptr_t *pred_stride,
                    void *const *rec, const intptr_t *rec_stride, int16_t *const *levels, hvb_tu_result *out);

/* Cooperative callers.  A thread that multiplexes many callers as fibers (integration/fiber_pool.cpp: the encoder's pool
 * threads, one per core, each running hundreds of CTU-row tasks) registers how a caller gives the thread up: in place of
 * the sleep, a waiting call invokes park(arg, done) until *done != 0 (park switches to the scheduler, which resumes the fiber
 * once *done is set); a session thread calls notify(arg) right after setting *done, so that a scheduler that went to sleep
 * with nothing runnable can be woken.  Per calling thread; NULL (or park == NULL) restores the blocking wait. */
typedef struct
{
    void (*park)(void *arg, const volatile int *done);
    void (*notify)(void *arg);
    void *arg;
} hvbenc_thread_hooks;
void hvbenc_set_thread_hooks(const hvbenc_thread_hooks *hooks);

/* counters since creation: tasks and batches per kind, device time.  Written as one JSON object into buf. */
int hvbenc_stats(hvbenc *enc, char *buf, size_t bytes);

#ifdef __cplusplus
}
#endif

#endif
