// integration/turing_hooks.cpp -- process-wide state of the hooks in turing_hooks.hpp: the one hvbenc session of the
// encoder process, created when the first picture arrives, and the per-thread memo of speculatively issued tasks.
#include "turing_hooks_state.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace hvbhooks {

namespace {
hvbenc *gSession = nullptr;
std::once_flag gOnce;

int envInt(const char *name, int fallback)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

void shutdown()
{
    if (!gSession) return;
    if (envInt("HVB_STATS", 0))
    {
        char buf[8192];
        if (!hvbenc_stats(gSession, buf, sizeof(buf))) fprintf(stderr, "hvbenc stats: %s\n", buf);
    }
    // The process is about to end: releasing gigabytes of device pictures and page-locked buffers one by one costs seconds the
    // driver's own teardown does not (HVB_FAST_EXIT=0 keeps the orderly release, e.g. under a leak checker).
    if (envInt("HVB_FAST_EXIT", 1)) return;
    hvbenc_destroy(gSession);
    gSession = nullptr;
}
} // namespace

bool on()
{
    static const bool value = envInt("HVB_BATCHED", 1) != 0;
    return value;
}

unsigned enabledMask()
{
    static const unsigned value = (unsigned)envInt("HVB_HOOKS", 0x3f);
    return value;
}

int intraMinLog2()
{
    static const int value = envInt("HVB_INTRA_MIN_LOG2", 2);
    return value;
}

int intraTuMinLog2()
{
    static const int value = envInt("HVB_INTRA_TU_MIN_LOG2", 4);
    return value;
}

int meMinArea()
{
    static const int value = envInt("HVB_ME_MIN_AREA", 0);
    return value;
}

int puMinArea()
{
    static const int value = envInt("HVB_PU_MIN_AREA", 0);
    return value;
}

int tuMinLog2()
{
    static const int value = envInt("HVB_TU_MIN_LOG2", 0);
    return value;
}

void fatal(const char *what, int rc)
{
    // there is no CPU fallback behind a failed device call: a wrong bitstream is worse than none
    fprintf(stderr, "turing_b200_batched: %s failed (%d): %s\n", what, rc, gSession ? hvbenc_last_error(gSession) : "no session");
    abort();
}

hvbenc *session(int bytesPerSample, int bitDepth, int width, int height)
{
    std::call_once(gOnce, [&] {
        const int rc = hvbenc_create(envInt("HVB_DEVICE", 0), bytesPerSample, bitDepth, width, height, envInt("HVB_POOL_PICTURES", 40), &gSession);
        if (rc)
        {
            fprintf(stderr, "turing_b200_batched: no usable B200 session (hvbenc_create: %d); there is no CPU fallback\n", rc);
            abort();
        }
        atexit(shutdown);
    });
    return gSession;
}

int pictureId(const void *key, bool fresh)
{
    int pic = -1;
    const int rc = hvbenc_picture(gSession, key, fresh, &pic);
    if (rc) fatal("hvbenc_picture", rc);
    return pic;
}

namespace {
thread_local CallerState *tlCaller = nullptr;

CallerState &caller()
{
    if (!tlCaller)
    {
        static thread_local CallerState *own = new CallerState(); // zero-initialised; lives as long as the thread may run hooks
        tlCaller = own;
    }
    return *tlCaller;
}
} // namespace

void setCallerState(CallerState *state)
{
    tlCaller = state;
}

TuMemo &tuMemo()
{
    return caller().tu;
}

Memo &memo()
{
    return caller().memo;
}

} // namespace hvbhooks
