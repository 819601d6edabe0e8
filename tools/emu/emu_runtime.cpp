// Runtime of the CPU emulation build (see include/cuda_runtime.h): cooperative fibers for the
// threads of one block, a fake synchronous CUDA runtime.  Test infrastructure only.
#include <setjmp.h>
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <chrono>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "cuda_runtime.h"

namespace emu {

namespace {

constexpr size_t kStackBytes = 256 * 1024;

// A fiber is born with makecontext/swapcontext (which set up its stack) and from then on
// switched with _setjmp/_longjmp: those do not save the signal mask, i.e. no system call per
// switch (swapcontext does one, and a kernel launch switches millions of times).
struct Fiber {
    ucontext_t uc;
    jmp_buf jb;
    ThreadCtx ctx;
    bool done = false;
    bool started = false;
    int warp = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    int alive = 0;
    int arrived = 0;
    unsigned gen = 0;
    std::vector<int> warpAlive, warpArrived;
    std::vector<unsigned> warpGen;
    std::vector<uint64_t> slots;     // 32 per warp
    std::vector<unsigned char> smem;
    const std::function<void()>* body = nullptr;
};

// per OS thread: the multi-rank tests run one rank per thread
thread_local Block g_block;
thread_local ucontext_t g_sched;
thread_local jmp_buf g_sched_jb;
thread_local Fiber* g_cur = nullptr;
thread_local std::vector<char*> g_stacks;
thread_local ThreadCtx g_hostCtx;

void release_if_complete() {
    Block& b = g_block;
    if (b.alive > 0 && b.arrived == b.alive) {
        b.arrived = 0;
        b.gen++;
    }
    for (size_t w = 0; w < b.warpAlive.size(); ++w)
        if (b.warpAlive[w] > 0 && b.warpArrived[w] == b.warpAlive[w]) {
            b.warpArrived[w] = 0;
            b.warpGen[w]++;
        }
}

void fiber_entry() {
    Fiber* f = g_cur;
    (*g_block.body)();
    f->done = true;
    g_block.alive--;
    g_block.warpAlive[f->warp]--;
    release_if_complete();   // threads that left do not hold the others at a barrier
    _longjmp(g_sched_jb, 1);
}

}  // namespace

ThreadCtx& cur() { return g_cur ? g_cur->ctx : g_hostCtx; }

void yield() {
    Fiber* f = g_cur;
    if (_setjmp(f->jb) == 0) _longjmp(g_sched_jb, 1);
}

void syncthreads() {
    Block& b = g_block;
    const unsigned gen = b.gen;
    b.arrived++;
    if (b.arrived == b.alive) {
        b.arrived = 0;
        b.gen++;
        return;
    }
    while (b.gen == gen) yield();
}

void syncwarp() {
    Block& b = g_block;
    const int w = g_cur->warp;
    const unsigned gen = b.warpGen[w];
    b.warpArrived[w]++;
    if (b.warpArrived[w] == b.warpAlive[w]) {
        b.warpArrived[w] = 0;
        b.warpGen[w]++;
        return;
    }
    while (b.warpGen[w] == gen) yield();
}

void* dyn_smem() { return g_block.smem.data(); }

uint64_t* warp_slots() { return g_block.slots.data() + 32 * g_cur->warp; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) {
        fprintf(stderr, "emu::launch: bad block size %d\n", nthreads);
        abort();
    }
    while ((int)g_stacks.size() < nthreads) g_stacks.push_back((char*)malloc(kStackBytes));
    Block& b = g_block;
    b.body = &body;
    const int nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                b.fibers.assign(nthreads, Fiber());
                b.alive = nthreads;
                b.arrived = 0;
                b.gen = 0;
                b.warpAlive.assign(nwarps, 0);
                b.warpArrived.assign(nwarps, 0);
                b.warpGen.assign(nwarps, 0);
                b.slots.assign(32 * (size_t)nwarps, 0);
                b.smem.assign(smem + 256, 0);
                for (int t = 0; t < nthreads; ++t) {
                    Fiber& f = b.fibers[t];
                    f.ctx.tid.x = t % block.x;
                    f.ctx.tid.y = (t / block.x) % block.y;
                    f.ctx.tid.z = t / (block.x * block.y);
                    f.ctx.bid.x = bx;
                    f.ctx.bid.y = by;
                    f.ctx.bid.z = bz;
                    f.ctx.bdim = block;
                    f.ctx.gdim = grid;
                    f.warp = t / 32;
                    b.warpAlive[f.warp]++;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = g_stacks[t];
                    f.uc.uc_stack.ss_size = kStackBytes;
                    f.uc.uc_link = nullptr;
                    makecontext(&f.uc, fiber_entry, 0);
                }
                // round robin until every thread of the block has returned.  ICQB_EMU_ORDER
                // changes the order in which the threads run between two synchronisation points
                // (reverse, or a fresh pseudo-random permutation every round): a missing barrier
                // that the ascending order happens to hide shows up as a different result.
                long idle = 0;
                static const char* orderEnv = getenv("ICQB_EMU_ORDER");
                const int orderMode = !orderEnv ? 0 : (orderEnv[0] == 'r' ? 1 : 2);
                std::vector<int> order(nthreads);
                for (int t = 0; t < nthreads; ++t) order[t] = orderMode == 1 ? nthreads - 1 - t : t;
                uint64_t rng = 0x9e3779b97f4a7c15ULL ^ (bx * 73856093u) ^ (by * 19349663u) ^ (bz * 83492791u);
                while (b.alive > 0) {
                    bool progressed = false;
                    if (orderMode == 2)
                        for (int t = nthreads - 1; t > 0; --t) {
                            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                            std::swap(order[t], order[(int)(rng % (uint64_t)(t + 1))]);
                        }
                    for (int ot = 0; ot < nthreads; ++ot) {
                        const int t = order[ot];
                        Fiber& f = b.fibers[t];
                        if (f.done) continue;
                        g_cur = &f;
                        if (_setjmp(g_sched_jb) == 0) {
                            if (!f.started) {
                                f.started = true;
                                setcontext(&f.uc);      // first entry: onto the fiber's stack
                            } else {
                                _longjmp(f.jb, 1);
                            }
                        }
                        g_cur = nullptr;
                        progressed = true;
                    }
                    if (!progressed) break;
                    if (++idle > 50000000L) {
                        fprintf(stderr, "emu::launch: block (%u,%u,%u) does not terminate (deadlock at a barrier?)\n",
                                bx, by, bz);
                        abort();
                    }
                }
            }
    b.body = nullptr;
}

}  // namespace emu

// ---- fake runtime ---------------------------------------------------------------------------
struct emuStream {
    int dummy;
};
struct emuEvent {
    std::chrono::steady_clock::time_point t;
};

static int g_num_sms = 4;

cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
    const char* e = getenv("ICQB_EMU_SMS");
    *v = e ? atoi(e) : g_num_sms;
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
// ICQB_EMU_GUARD=1: every "device" allocation ends exactly (to 16 bytes) at an inaccessible
// page, so a kernel or copy that runs past the end of a buffer faults at once instead of reading
// the allocator's padding -- the emulation's stand-in for compute-sanitizer's memcheck.
// ICQB_EMU_GUARD=2 puts the inaccessible page BEFORE the buffer instead (underruns).
static int guard_mode() {
    static const int g = getenv("ICQB_EMU_GUARD") ? atoi(getenv("ICQB_EMU_GUARD")) : 0;
    return g;
}
static std::mutex g_guard_mu;
static std::unordered_map<void*, std::pair<void*, size_t>> g_guarded;   // user pointer -> mapping

cudaError_t emuMalloc(void** p, size_t bytes) {
    if (guard_mode()) {
        const size_t page = 4096;
        const size_t b16 = (bytes + 15) / 16 * 16;
        const size_t body = (b16 + page - 1) / page * page;
        char* base = (char*)mmap(nullptr, body + page, PROT_READ | PROT_WRITE,
                                 MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (base == MAP_FAILED) return cudaErrorInvalidValue;
        void* q;
        if (guard_mode() == 2) {
            mprotect(base, page, PROT_NONE);
            memset(base + page, 0xff, body);
            q = base + page;
        } else {
            memset(base, 0xff, body);
            mprotect(base + body, page, PROT_NONE);
            q = base + body - b16;
        }
        std::lock_guard<std::mutex> lock(g_guard_mu);
        g_guarded[q] = std::make_pair((void*)base, body + page);
        *p = q;
        return cudaSuccess;
    }
    // poison fresh memory with NaN patterns: reads of never-written "device" memory show up
    size_t n = (bytes + 63) / 64 * 64 + 64;
    void* q = aligned_alloc(256, (n + 255) / 256 * 256);
    if (!q) return cudaErrorInvalidValue;
    memset(q, 0xff, n);
    *p = q;
    return cudaSuccess;
}
static void emu_release(void* p) {
    if (!p) return;
    if (guard_mode()) {
        std::lock_guard<std::mutex> lock(g_guard_mu);
        auto it = g_guarded.find(p);
        if (it != g_guarded.end()) {
            munmap(it->second.first, it->second.second);
            g_guarded.erase(it);
            return;
        }
    }
    free(p);
}
cudaError_t cudaFree(void* p) { emu_release(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { emu_release(p); return cudaSuccess; }
// for the tests: a guarded (or plain) buffer to stand in for a caller-owned device array
extern "C" void* emu_alloc(size_t bytes) {
    void* p = nullptr;
    return emuMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
}
extern "C" void emu_free(void* p) { emu_release(p); }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
    memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t) {
    return cudaMemcpy(dst, src, bytes, k);
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                              size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r)
        memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t) {
    memset(p, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int value, size_t bytes) {
    memset(p, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

// ---- in-process stand-in for NCCL: one rank per OS thread ----------------------------------------
// (bound by csrc/comm.cu under ICQB_EMU instead of dlopen("libnccl.so.2"))
#include <condition_variable>
#include <map>
#include <mutex>

namespace {

struct World {
    int nranks = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned gen = 0;
    std::vector<const void*> send;

    void barrier() {
        std::unique_lock<std::mutex> lock(m);
        const unsigned g = gen;
        if (++arrived == nranks) {
            arrived = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lock, [&] { return gen != g; });
        }
    }
};

struct FakeComm {
    World* world;
    int rank;
};

std::mutex g_worlds_mutex;
std::map<long long, World*> g_worlds;
long long g_next_id = 1;

}  // namespace

extern "C" {

int emu_ncclGetUniqueId(void* id128) {
    std::lock_guard<std::mutex> lock(g_worlds_mutex);
    memset(id128, 0, 128);
    const long long id = g_next_id++;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int emu_ncclCommInitRank(void** comm, int nranks, const void* id128, int rank) {
    long long id = 0;
    memcpy(&id, id128, sizeof(id));
    World* w;
    {
        std::lock_guard<std::mutex> lock(g_worlds_mutex);
        World*& slot = g_worlds[id];
        if (!slot) {
            slot = new World();
            slot->nranks = nranks;
            slot->send.assign(nranks, nullptr);
        }
        w = slot;
    }
    *comm = new FakeComm{w, rank};
    w->barrier();
    return 0;
}

int emu_ncclCommDestroy(void* comm) {
    delete static_cast<FakeComm*>(comm);
    return 0;
}

// sum of doubles in rank order, identical on every rank
int emu_ncclAllReduceSumF64(const void* send, void* recv, size_t count, void* comm) {
    FakeComm* c = static_cast<FakeComm*>(comm);
    World* w = c->world;
    w->send[c->rank] = send;
    w->barrier();
    std::vector<double> tmp(count, 0.0);
    for (int r = 0; r < w->nranks; ++r) {
        const double* s = static_cast<const double*>(w->send[r]);
        for (size_t k = 0; k < count; ++k) tmp[k] += s[k];
    }
    w->barrier();
    memcpy(recv, tmp.data(), sizeof(double) * count);
    w->barrier();
    return 0;
}

int emu_ncclAllReduceMaxF64(const void* send, void* recv, size_t count, void* comm) {
    FakeComm* c = static_cast<FakeComm*>(comm);
    World* w = c->world;
    w->send[c->rank] = send;
    w->barrier();
    std::vector<double> tmp(static_cast<const double*>(w->send[0]), static_cast<const double*>(w->send[0]) + count);
    for (int r = 1; r < w->nranks; ++r) {
        const double* s = static_cast<const double*>(w->send[r]);
        for (size_t k = 0; k < count; ++k) tmp[k] = s[k] > tmp[k] ? s[k] : tmp[k];
    }
    w->barrier();
    memcpy(recv, tmp.data(), sizeof(double) * count);
    w->barrier();
    return 0;
}

int emu_ncclBroadcastF64(const void* send, void* recv, size_t count, int root, void* comm) {
    FakeComm* c = static_cast<FakeComm*>(comm);
    World* w = c->world;
    w->send[c->rank] = send;
    w->barrier();
    std::vector<double> tmp(static_cast<const double*>(w->send[root]),
                            static_cast<const double*>(w->send[root]) + count);
    w->barrier();
    memcpy(recv, tmp.data(), sizeof(double) * count);
    w->barrier();
    return 0;
}

int emu_ncclAllGatherF64(const void* send, void* recv, size_t count, void* comm) {
    FakeComm* c = static_cast<FakeComm*>(comm);
    World* w = c->world;
    w->send[c->rank] = send;
    w->barrier();
    std::vector<double> tmp(count * w->nranks);
    for (int r = 0; r < w->nranks; ++r) memcpy(tmp.data() + r * count, w->send[r], sizeof(double) * count);
    w->barrier();
    memcpy(recv, tmp.data(), sizeof(double) * count * w->nranks);
    w->barrier();
    return 0;
}

}  // extern "C"
