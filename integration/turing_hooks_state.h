// integration/turing_hooks_state.h -- what turing_hooks.hpp (compiled into the reference encoder's translation units) and
// turing_hooks.cpp (the process-wide state) share; free of reference types.
#ifndef INCLUDED_turing_hooks_state_h
#define INCLUDED_turing_hooks_state_h

#include "hvb_encoder.h"
#include <cstring>

namespace hvbhooks {

bool on();                                                        // HVB_BATCHED=0 in the environment switches the hooks off
hvbenc *session(int bytesPerSample, int bitDepth, int width, int height); // created on first use
int pictureId(const void *key, bool fresh);
void fatal(const char *what, int rc);
int intraMinLog2();                                              // HVB_INTRA_MIN_LOG2 (default 2): smallest partition whose sweep goes to the device
unsigned enabledMask();                                           // HVB_HOOKS: bit 0 me, 1 bi, 2 pu cost, 3 intra sweep, 4 inter tu, 5 intra tu
int intraTuMinLog2();                                            // HVB_INTRA_TU_MIN_LOG2 (default 4): smallest intra transform block whose chain goes to the device
// Size thresholds: a block below them stays with the reference's own body (its CPU havoc tables), because a hand-over costs
// more host time than the block's arithmetic does there; results are bit-exact either way, so the bitstream cannot tell.
int meMinArea();                                                 // HVB_ME_MIN_AREA: smallest PU (w*h) whose uni / bi search goes to the device
int puMinArea();                                                 // HVB_PU_MIN_AREA: smallest PU (w*h) whose prediction + SATD goes to the device
int tuMinLog2();                                                 // HVB_TU_MIN_LOG2: smallest inter CU (log2 of the transform tree's root) whose blocks go to the device

struct Memo // per-thread results of tasks issued ahead of the reference's control flow
{
    static const int kMe = 4, kPu = 16;
    hvb_me_task meTask[kMe];
    hvb_me_result meResult[kMe];
    int nMe;
    hvb_pu_cost_task puTask[kPu];
    int32_t puResult[kPu][3];
    int nPu;
    void clear() { nMe = nPu = 0; }
};
Memo &memo();

struct TuBlock // what the device returns for one transform block
{
    int x0, y0, cIdx, log2n;
    uint32_t ssd, ssdPred;
    int cbf;
    int16_t levels[32 * 32];
};

struct TuMemo // per-thread: the blocks of the CU being reconstructed, issued together ahead of the tree walk
{
    static const int kBlocks = 24;
    TuBlock block[kBlocks];
    int n;
    bool active; // the CU being reconstructed has its blocks on the device
};
TuMemo &tuMemo();

// The memos belong to a CALLER (one CTU row being encoded), not to a thread: integration/fiber_pool.cpp runs many rows per
// pool thread and switches the caller's state with the fiber.  A thread that runs its tasks directly uses its own.
struct CallerState
{
    Memo memo;
    TuMemo tu;
    void clear()
    {
        memo.clear();
        tu.n = 0;
        tu.active = false;
    }
};
void setCallerState(CallerState *state); // nullptr: back to the calling thread's own



} // namespace hvbhooks

#endif
