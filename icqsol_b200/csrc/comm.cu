// NCCL plumbing for the row-sharded solve: the CG iterate is all-gathered and the
// dot products all-reduced (the reference has no multi-process code at all;
// SURVEY.md section 8e).  NCCL is bound at run time with dlopen so that the library
// resolves to the libnccl the host process already carries (PyTorch's bundled
// 2.28.9 when driven from the Python layer) and has no link-time dependency.
#include <dlfcn.h>
#include <string.h>

#include "icqb_internal.cuh"

namespace icqb {

namespace {

// the slice of the NCCL ABI used here (stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclFloat64 = 8;
constexpr int ncclSum = 0;
constexpr int ncclMax = 2;

struct Api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

#ifdef ICQB_EMU   // CPU emulation build of the test suite (tools/emu): ranks are threads of one process
extern "C" {
int emu_ncclGetUniqueId(void* id128);
int emu_ncclCommInitRank(void** comm, int nranks, const void* id128, int rank);
int emu_ncclCommDestroy(void* comm);
int emu_ncclAllReduceSumF64(const void* send, void* recv, size_t count, void* comm);
int emu_ncclAllGatherF64(const void* send, void* recv, size_t count, void* comm);
int emu_ncclAllReduceMaxF64(const void* send, void* recv, size_t count, void* comm);
int emu_ncclBroadcastF64(const void* send, void* recv, size_t count, int root, void* comm);
}
Api& api() {
    static Api a;
    if (a.handle) return a;
    a.handle = &a;
    a.GetUniqueId = [](ncclUniqueId* id) -> ncclResult_t { return emu_ncclGetUniqueId(id); };
    a.CommInitRank = [](ncclComm_t* c, int n, ncclUniqueId id, int r) -> ncclResult_t {
        return emu_ncclCommInitRank(reinterpret_cast<void**>(c), n, &id, r);
    };
    a.CommDestroy = [](ncclComm_t c) -> ncclResult_t { return emu_ncclCommDestroy(c); };
    a.AllGather = [](const void* s, void* r, size_t n, int, ncclComm_t c, cudaStream_t) -> ncclResult_t {
        return emu_ncclAllGatherF64(s, r, n, c);
    };
    a.AllReduce = [](const void* s, void* r, size_t n, int, int op, ncclComm_t c,
                     cudaStream_t) -> ncclResult_t {
        return op == ncclMax ? emu_ncclAllReduceMaxF64(s, r, n, c) : emu_ncclAllReduceSumF64(s, r, n, c);
    };
    a.Broadcast = [](const void* s, void* r, size_t n, int, int root, ncclComm_t c,
                     cudaStream_t) -> ncclResult_t { return emu_ncclBroadcastF64(s, r, n, root, c); };
    a.GetErrorString = [](ncclResult_t) -> const char* { return "emulated NCCL error"; };
    return a;
}
#else
Api& api() {
    static Api a;
    if (a.handle || !a.error.empty()) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) {
        a.error = std::string("cannot load libnccl: ") + dlerror();
        return a;
    }
#define ICQB_SYM(field, name)                                              \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name)); \
    if (!a.field) a.error = std::string("libnccl lacks ") + name;
    ICQB_SYM(GetUniqueId, "ncclGetUniqueId")
    ICQB_SYM(CommInitRank, "ncclCommInitRank")
    ICQB_SYM(CommDestroy, "ncclCommDestroy")
    ICQB_SYM(AllGather, "ncclAllGather")
    ICQB_SYM(AllReduce, "ncclAllReduce")
    ICQB_SYM(Broadcast, "ncclBroadcast")
    ICQB_SYM(GetErrorString, "ncclGetErrorString")
#undef ICQB_SYM
    return a;
}
#endif

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int refs = 1;   // contexts sharing this communicator (icqb_comm_share)
    // optional peer exchange (SymvPeer, icqb_internal.cuh)
    void* xbuf = nullptr;               // this rank's exchange buffer
    void* xpeer[kMaxPeers] = {};        // every rank's buffer as mapped here (own: xbuf)
    long long slotStride = 0;
    bool xattached = false;
    unsigned long long seq = 0;
};

constexpr size_t kPeerHeaderBytes = 256;

static size_t peer_bytes(long long max_n, int nranks) {
    return kPeerHeaderBytes + sizeof(double) * 2 * (size_t)nranks * (size_t)(max_n + 8);
}

static void peer_destroy(Comm* c) {
    for (int h = 0; h < c->nranks && h < kMaxPeers; ++h)
        if (c->xpeer[h] && c->xpeer[h] != c->xbuf) cudaIpcCloseMemHandle(c->xpeer[h]);
    if (c->xbuf) cudaFree(c->xbuf);
    c->xbuf = nullptr;
    c->xattached = false;
    for (int h = 0; h < kMaxPeers; ++h) c->xpeer[h] = nullptr;
}

bool comm_peer_next(icqb_ctx* ctx, long long n, SymvPeer* out) {
    Comm* c = ctx->comm;
    if (!c || !c->xattached || n + 1 > c->slotStride) return false;
    out->nranks = c->nranks;
    out->rank = c->rank;
    out->seq = ++c->seq;
    out->parity = (int)(out->seq & 1);
    out->slotStride = c->slotStride;
    for (int h = 0; h < c->nranks; ++h) {
        char* base = static_cast<char*>(c->xpeer[h]);
        out->flags[h] = reinterpret_cast<unsigned long long*>(base);
        out->slots[h] = reinterpret_cast<double*>(base + kPeerHeaderBytes);
    }
    return true;
}

int comm_nranks(const icqb_ctx* ctx) { return ctx->comm ? ctx->comm->nranks : 1; }
int comm_rank(const icqb_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }

static int nccl_check(icqb_ctx* ctx, ncclResult_t r, const char* what) {
    if (r == 0) return ICQB_OK;
    return set_error(ctx, ICQB_ECOMM, std::string(what) + ": " + api().GetErrorString(r));
}

int comm_allgather(icqb_ctx* ctx, const double* send, double* recv, int64_t count_per_rank,
                   cudaStream_t stream) {
    if (comm_nranks(ctx) == 1) {
        if (send != recv)
            return check_cuda(ctx,
                              cudaMemcpyAsync(recv, send, sizeof(double) * count_per_rank,
                                              cudaMemcpyDeviceToDevice, stream),
                              "allgather(self)");
        return ICQB_OK;
    }
    return nccl_check(ctx,
                      api().AllGather(send, recv, (size_t)count_per_rank, ncclFloat64,
                                      ctx->comm->comm, stream),
                      "ncclAllGather");
}

int comm_allreduce_sum(icqb_ctx* ctx, double* buf, int64_t count, cudaStream_t stream) {
    if (comm_nranks(ctx) == 1) return ICQB_OK;
    return nccl_check(ctx,
                      api().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum,
                                      ctx->comm->comm, stream),
                      "ncclAllReduce");
}

int comm_allreduce_max(icqb_ctx* ctx, double* buf, int64_t count, cudaStream_t stream) {
    if (comm_nranks(ctx) == 1) return ICQB_OK;
    return nccl_check(ctx,
                      api().AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclMax,
                                      ctx->comm->comm, stream),
                      "ncclAllReduce(max)");
}

int comm_broadcast(icqb_ctx* ctx, double* buf, int64_t count, int root, cudaStream_t stream) {
    if (comm_nranks(ctx) == 1 || count <= 0) return ICQB_OK;
    return nccl_check(ctx,
                      api().Broadcast(buf, buf, (size_t)count, ncclFloat64, root, ctx->comm->comm,
                                      stream),
                      "ncclBroadcast");
}

void comm_destroy(icqb_ctx* ctx) {
    if (!ctx->comm) return;
    if (--ctx->comm->refs == 0) {
        peer_destroy(ctx->comm);
        if (ctx->comm->comm) api().CommDestroy(ctx->comm->comm);
        delete ctx->comm;
    }
    ctx->comm = nullptr;
}

}  // namespace icqb

using namespace icqb;

extern "C" {

int icqb_comm_unique_id(void* unique_id_128) {
    if (!unique_id_128) return ICQB_EARG;
    Api& a = api();
    if (!a.error.empty()) return set_error(nullptr, ICQB_ECOMM, a.error);
    ncclUniqueId id;
    ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return set_error(nullptr, ICQB_ECOMM, std::string("ncclGetUniqueId: ") + a.GetErrorString(r));
    memcpy(unique_id_128, &id, sizeof(id));
    return ICQB_OK;
}

int icqb_comm_init(icqb_ctx* ctx, const void* unique_id_128, int rank, int nranks) {
    if (!ctx || !unique_id_128 || nranks < 1 || rank < 0 || rank >= nranks)
        return set_error(ctx, ICQB_EARG, "icqb_comm_init: bad arguments");
    comm_destroy(ctx);
    if (nranks == 1) return ICQB_OK;
    Api& a = api();
    if (!a.error.empty()) return set_error(ctx, ICQB_ECOMM, a.error);
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    Comm* c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    ncclResult_t r = a.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
        delete c;
        return set_error(ctx, ICQB_ECOMM, std::string("ncclCommInitRank: ") + a.GetErrorString(r));
    }
    ctx->comm = c;
    return ICQB_OK;
}

int icqb_comm_share(icqb_ctx* ctx, icqb_ctx* owner) {
    if (!ctx || !owner) return ICQB_EARG;
    if (ctx->device != owner->device)
        return set_error(ctx, ICQB_EARG, "icqb_comm_share: contexts live on different devices");
    comm_destroy(ctx);
    if (owner->comm) {
        ctx->comm = owner->comm;
        ctx->comm->refs += 1;
    }
    return ICQB_OK;
}

int icqb_peer_create(icqb_ctx* ctx, int64_t max_n, void* handle64) {
    if (!ctx || !handle64 || max_n <= 0) return set_error(ctx, ICQB_EARG, "icqb_peer_create: bad arguments");
    Comm* c = ctx->comm;
    if (!c) return set_error(ctx, ICQB_ESTATE, "icqb_peer_create: no communicator (single rank)");
    if (c->nranks > kMaxPeers) return set_error(ctx, ICQB_EARG, "icqb_peer_create: at most 16 ranks");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    peer_destroy(c);
    const size_t bytes = peer_bytes(max_n, c->nranks);
    ICQB_CUDA(ctx, cudaMalloc(&c->xbuf, bytes));
    ICQB_CUDA(ctx, cudaMemset(c->xbuf, 0, bytes));
    c->slotStride = max_n + 8;
    c->seq = 0;
    cudaIpcMemHandle_t h;
    ICQB_CUDA(ctx, cudaIpcGetMemHandle(&h, c->xbuf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    return ICQB_OK;
}

int icqb_peer_attach(icqb_ctx* ctx, const void* handles) {
    if (!ctx || !handles) return set_error(ctx, ICQB_EARG, "icqb_peer_attach: bad arguments");
    Comm* c = ctx->comm;
    if (!c || !c->xbuf) return set_error(ctx, ICQB_ESTATE, "icqb_peer_attach: call icqb_peer_create first");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int h = 0; h < c->nranks; ++h) {
        if (h == c->rank) {
            c->xpeer[h] = c->xbuf;
            continue;
        }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, static_cast<const char*>(handles) + 64 * h, 64);
        ICQB_CUDA(ctx, cudaIpcOpenMemHandle(&c->xpeer[h], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    c->xattached = true;
    return ICQB_OK;
}

int icqb_comm_rank(const icqb_ctx* ctx, int* rank, int* nranks) {
    if (!ctx) return ICQB_EARG;
    if (rank) *rank = comm_rank(ctx);
    if (nranks) *nranks = comm_nranks(ctx);
    return ICQB_OK;
}

}  // extern "C"
