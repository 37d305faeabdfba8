// integration/fiber_pool.cpp -- the encoder's thread pool (interface of turing/ThreadPool.h, unchanged) with tasks run as
// fibers, linked into turing_b200_batched / turing_b200_segments in place of turing/ThreadPool.cpp.
//
// Why: the reference's pool threads (turing/ThreadPool.cpp:87-103) run a CTU row each (turing/TaskEncodeSubstream.cpp:150-215)
// and call the pixel primitives directly.  On the batched path every such call is a hand-over to the device that takes a
// few hundred microseconds to come back.  With a thread per row the hand-over is a futex sleep and a wake-up (more host
// time than the AVX2 primitive it replaces), and the number of hand-overs in flight -- which is what the batch size and
// hence the device's efficiency depend on -- is capped by the thread count.  Here a pool thread is a scheduler: a task's
// run() executes on its own stack; when it waits for the device (hvbenc_set_thread_hooks: park) the scheduler switches to
// another runnable row or starts the next unblocked task, and resumes the row when its answer has arrived.  One pool thread
// per core carries hundreds of rows; a hand-over costs two stack switches.
//
// Semantics kept from the reference: a task is taken from the backlog when blocked() is false (scan in backlog order under
// the pool mutex), run() returning true puts it back at the front, a null task ends a worker, nudge() wakes sleeping workers.
// A fiber stays on the thread that started it (thread-local state of the hooks and of libhvb is per scheduler thread; the
// per-caller memo of turing_hooks is switched with the fiber).  No lock is held across a hand-over: the hooks sit outside
// the regions the tasks guard with the pool mutex.
//
// HVB_FIBERS=<n>: rows in flight per pool thread (default 48); 0 = the reference's thread-per-task loop.
#include "ThreadPool.h"

#include "turing_hooks_state.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sys/mman.h>

namespace {

#if !defined(__x86_64__)
#error "fiber_pool.cpp: the stack switch is written for x86-64"
#endif

// six callee-saved registers and the stack pointer: all the state a cooperative switch between two C++ frames needs
__attribute__((naked, noinline)) void switchStack(void **saveSp, void *loadSp)
{
    asm volatile("pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
                 "movq %rsp, (%rdi)\n movq %rsi, %rsp\n"
                 "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n ret\n");
}

// nudge() counters, one per pool (ThreadPool.h is the reference's: no room for a member).  A worker rescans the backlog only
// when its pool was nudged or one of its own tasks returned; a hand-over -- tens of thousands per picture -- must not cost
// a walk over the backlog under the pool mutex.
struct PoolCounter
{
    std::atomic<ThreadPool *> pool{nullptr};
    std::atomic<uint64_t> nudges{0};
};
constexpr int kMaxPools = 256;
PoolCounter gCounters[kMaxPools];
thread_local bool tlQuietNudge = false; // notify(): wake sleepers without announcing a change of the backlog

std::atomic<uint64_t> &counterOf(ThreadPool *pool)
{
    for (int i = 0; i < kMaxPools; ++i)
    {
        ThreadPool *seen = gCounters[i].pool.load(std::memory_order_acquire);
        if (seen == pool) return gCounters[i].nudges;
        if (!seen)
        {
            ThreadPool *expected = nullptr;
            if (gCounters[i].pool.compare_exchange_strong(expected, pool) || expected == pool) return gCounters[i].nudges;
        }
    }
    return gCounters[kMaxPools - 1].nudges; // more pools than slots: they share a counter (spurious rescans only)
}

void releaseCounter(ThreadPool *pool)
{
    for (int i = 0; i < kMaxPools; ++i)
        if (gCounters[i].pool.load(std::memory_order_acquire) == pool) gCounters[i].pool.store(nullptr, std::memory_order_release);
}

struct Fiber
{
    void *sp = nullptr;
    char *stack = nullptr;
    size_t stackBytes = 0;
    ThreadPool::Task *task = nullptr;
    const volatile int *waitingFor = nullptr;
    bool finished = false, blockedResult = false;
    hvbhooks::CallerState state; // the hooks' memo of the row this fiber runs
};

struct Worker
{
    ThreadPool *pool = nullptr;
    void *schedulerSp = nullptr;
    Fiber *current = nullptr;
    std::vector<Fiber *> parked, spare;
    int live = 0;
    std::atomic<int> ready{0}, sleeping{0};
};

thread_local Worker *tlWorker = nullptr;

int envInt(const char *name, int fallback)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

void fiberMain()
{
    for (;;)
    {
        Worker *w = tlWorker;
        Fiber *f = w->current;
        f->blockedResult = f->task->run();
        f->finished = true;
        switchStack(&f->sp, tlWorker->schedulerSp);
    }
}

Fiber *newFiber()
{
    static const size_t bytes = (size_t)envInt("HVB_FIBER_STACK_KB", 8192) * 1024;
    Fiber *f = new Fiber();
    const size_t guard = 4096;
    void *p = mmap(nullptr, bytes + guard, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
    if (p == MAP_FAILED)
    {
        perror("fiber_pool: mmap of a fiber stack");
        abort();
    }
    mprotect(p, guard, PROT_NONE);
    f->stack = static_cast<char *>(p);
    f->stackBytes = bytes + guard;
    // the first switch pops six registers and returns into fiberMain with the stack as after a call
    uintptr_t top = (reinterpret_cast<uintptr_t>(f->stack) + f->stackBytes) & ~uintptr_t(15);
    void **sp = reinterpret_cast<void **>(top - 16);
    *sp = reinterpret_cast<void *>(&fiberMain);
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    f->sp = sp;
    return f;
}

void freeFiber(Fiber *f)
{
    munmap(f->stack, f->stackBytes);
    delete f;
}

// hvbenc's hook: the calling fiber waits for *done
void park(void *arg, const volatile int *done)
{
    Worker *w = static_cast<Worker *>(arg);
    Fiber *f = w->current;
    if (!f) return; // called from the scheduler's own stack (not a task): the caller's loop polls
    f->waitingFor = done;
    switchStack(&f->sp, w->schedulerSp);
}

// hvbenc's hook, on a session thread: an answer for one of this worker's fibers is there
void notify(void *arg)
{
    Worker *w = static_cast<Worker *>(arg);
    w->ready.store(1, std::memory_order_seq_cst);
    if (w->sleeping.load(std::memory_order_seq_cst))
    {
        // No pass through the pool mutex here (tasks and scanning workers hold it for long stretches, and this runs on the
        // dispatcher's time): a wake-up that falls between the worker's last look at `ready` and its wait is lost, which the
        // wait's short time-out bounds.
        tlQuietNudge = true;
        w->pool->nudge();
        tlQuietNudge = false;
    }
}

} // namespace

ThreadPool::ThreadPool(int n)
{
    if (n == 0) n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    for (int i = 0; i < n; ++i) this->threads.push_back(std::thread(&ThreadPool::worker, this));
}

ThreadPool::~ThreadPool()
{
    {
        std::unique_lock<std::mutex> lock(this->poolMutex);
        for (unsigned i = 0; i < this->size(); ++i) this->backlog.push_back(0);
    }
    this->nudge();
    for (auto &thread : this->threads) thread.join();
    releaseCounter(this);
}

void ThreadPool::add(Task &task)
{
    std::unique_lock<std::mutex> lock(this->poolMutex);
    this->backlog.push_back(&task);
    this->nudge();
}

// blocking form, as the reference's: the first task of the backlog that can run
ThreadPool::Task *ThreadPool::getNextTask()
{
    std::unique_lock<std::mutex> lock(this->poolMutex);
    while (true)
    {
        for (auto i = this->backlog.begin(); i != this->backlog.end(); ++i)
        {
            ThreadPool::Task *task = *i;
            if (!task || !task->blocked())
            {
                this->backlog.erase(i);
                return task;
            }
        }
        this->taskAvailable.wait(lock);
    }
}

void ThreadPool::worker()
{
    const int capacity = envInt("HVB_FIBERS", 48);
    if (capacity <= 0 || !hvbhooks::on())
    {
        // the reference's loop: a task per thread, hand-overs sleep
        while (true)
        {
            Task *task = getNextTask();
            if (!task) break;
            if (task->run())
            {
                std::unique_lock<std::mutex> lock(this->poolMutex);
                this->backlog.push_front(task);
            }
        }
        return;
    }

    Worker &w = *new Worker(); // never freed: a session thread may still be inside notify() when the last answer is consumed
    w.pool = this;
    tlWorker = &w;
    const hvbenc_thread_hooks hooks = {park, notify, &w};
    hvbenc_set_thread_hooks(&hooks);
    const auto spinFor = std::chrono::microseconds(envInt("HVB_FIBER_SPIN_US", 0));
    std::atomic<uint64_t> &nudgeCounter = counterOf(this);

    bool rescan = true;
    // run fiber f until it parks or its task returns
    auto resume = [&](Fiber *f) {
        f->waitingFor = nullptr;
        w.current = f;
        hvbhooks::setCallerState(&f->state);
        switchStack(&w.schedulerSp, f->sp);
        hvbhooks::setCallerState(nullptr);
        w.current = nullptr;
        if (!f->finished)
        {
            w.parked.push_back(f);
            return;
        }
        --w.live;
        rescan = true; // a task of this worker returned: the reference's loop would look at the backlog now
        if (f->blockedResult)
        {
            std::unique_lock<std::mutex> lock(this->poolMutex);
            this->backlog.push_front(f->task);
        }
        f->task = nullptr;
        w.spare.push_back(f);
    };

    bool exiting = false;
    uint64_t seenNudges = ~uint64_t(0);
    for (;;)
    {
        bool progressed = false;
        // 1. rows whose answer has arrived, oldest first
        w.ready.store(0, std::memory_order_seq_cst);
        for (size_t i = 0; i < w.parked.size();)
        {
            Fiber *f = w.parked[i];
            if (*f->waitingFor)
            {
                std::atomic_thread_fence(std::memory_order_acquire);
                w.parked.erase(w.parked.begin() + i);
                resume(f);
                progressed = true;
            }
            else
                ++i;
        }
        // 2. the next task that can run, while there is room.  With rows parked the scan must not block; it is repeated
        // when something may have changed: a nudge, or one of this worker's own tasks returned or advanced.
        if (!exiting && w.live < capacity)
        {
            Task *task = nullptr;
            bool got = false;
            if (w.live == 0)
            {
                task = getNextTask();
                got = true;
            }
            else
            {
                const uint64_t nudges = nudgeCounter.load(std::memory_order_acquire);
                if (rescan || nudges != seenNudges)
                {
                    std::unique_lock<std::mutex> lock(this->poolMutex);
                    for (auto i = this->backlog.begin(); i != this->backlog.end(); ++i)
                    {
                        Task *t = *i;
                        if (!t || !t->blocked())
                        {
                            this->backlog.erase(i);
                            task = t;
                            got = true;
                            break;
                        }
                    }
                    if (!got)
                    {
                        seenNudges = nudges;
                        rescan = false;
                    }
                }
            }
            if (got)
            {
                progressed = true;
                if (!task)
                    exiting = true;
                else
                {
                    Fiber *f;
                    if (w.spare.empty())
                        f = newFiber();
                    else
                    {
                        f = w.spare.back();
                        w.spare.pop_back();
                    }
                    f->task = task;
                    f->finished = false;
                    f->state.clear();
                    ++w.live;
                    rescan = true; // more tasks may be runnable: keep taking them while there is room
                    resume(f);
                }
            }
        }
        if (exiting && w.live == 0) break;
        if (progressed) continue;
        // 3. nothing runnable: every row of this worker waits for the device.  Spin briefly (answers arrive every few tens
        // of microseconds under load), then sleep on the pool's condition variable; notify() and nudge() both end the sleep.
        const auto t0 = std::chrono::steady_clock::now();
        bool found = false;
        while (!found && std::chrono::steady_clock::now() - t0 < spinFor)
        {
            if (w.ready.load(std::memory_order_relaxed) || nudgeCounter.load(std::memory_order_relaxed) != seenNudges) found = true;
            else
                __builtin_ia32_pause();
        }
        if (found) continue;
        {
            std::unique_lock<std::mutex> lock(this->poolMutex);
            w.sleeping.store(1, std::memory_order_seq_cst);
            if (!w.ready.load(std::memory_order_seq_cst) && nudgeCounter.load(std::memory_order_acquire) == seenNudges)
                this->taskAvailable.wait_for(lock, std::chrono::microseconds(250));
            w.sleeping.store(0, std::memory_order_seq_cst);
        }
    }

    hvbenc_set_thread_hooks(nullptr);
    tlWorker = nullptr;
    for (Fiber *f : w.spare) freeFiber(f);
}

void ThreadPool::lock()
{
    this->poolMutex.lock();
}

void ThreadPool::unlock()
{
    this->poolMutex.unlock();
}

std::mutex &ThreadPool::mutex()
{
    return this->poolMutex;
}

void ThreadPool::nudge()
{
    if (!tlQuietNudge) counterOf(this).fetch_add(1, std::memory_order_release);
    this->taskAvailable.notify_all();
}

size_t ThreadPool::size() const
{
    return this->threads.size();
}
