// Context lifecycle, mesh upload and per-triangle precomputation.
//
// Replaces the set-up part of computeOffDiagonalTerms
// (bem/icqLaplaceMatrices.cpp:40-71: quadrature handle, connectivity copy) and
// hoists what the reference recomputes per pair -- the source quadrature points
// (bem/icqQuadrature.cpp:237-240, 16x per pair) and getArea
// (bem/icqLaplaceMatrices.cpp:9-30, once per pair) -- into one O(N) kernel.
#include <stdio.h>

#include <stdlib.h>

#include <map>
#include <mutex>

#include "icqb_internal.cuh"
#include "quad_tables.cuh"

namespace icqb {

static std::string g_create_error;
static std::mutex g_mutex;

int set_error(icqb_ctx* ctx, int code, const std::string& msg) {
    if (ctx)
        ctx->err = msg;
    else {
        std::lock_guard<std::mutex> lock(g_mutex);
        g_create_error = msg;
    }
    return code;
}

int check_cuda(icqb_ctx* ctx, cudaError_t e, const char* what) {
    // ICQB_DEBUG_SYNC=1 (debugging aid): wait for the device after every checked call, so that
    // an asynchronous fault is reported by the launch that caused it, not by a later call
    static const bool debugSync = [] {
        const char* v = getenv("ICQB_DEBUG_SYNC");
        return v && v[0] == '1';
    }();
    if (e == cudaSuccess && debugSync && ctx) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) return ICQB_OK;
    return set_error(ctx, ICQB_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int lower_layout(icqb_ctx* ctx, LowerLayout* L) {
    if ((ctx->r1 > ctx->r0 && (ctx->r0 % kLowerTile) != 0) ||
        (ctx->q1 > ctx->q0 && (ctx->q0 % kLowerTile) != 0))
        return set_error(ctx, ICQB_EARG, "lower storage: row blocks must start on multiples of 128");
    L->r0 = ctx->r0; L->r1 = ctx->r1; L->q0 = ctx->q0; L->q1 = ctx->q1;
    L->nTiles0 = (int)((L->r1 - L->r0 + kLowerTile - 1) / kLowerTile);
    L->nTiles1 = (int)((L->q1 - L->q0 + kLowerTile - 1) / kLowerTile);
    const long long I0 = L->r0 / kLowerTile, k0 = L->nTiles0;
    L->tiles0 = k0 * (I0 + 1) + k0 * (k0 - 1) / 2;
    return ICQB_OK;
}

// Process-wide cache of freed device blocks, keyed by (device, size).  A process that builds one
// solver after the other on the same mesh size (bench.py's end-to-end loop, a parameter sweep)
// asks for the same ~20 buffers again and again; cudaMalloc / cudaFree cost 0.1-1 ms each (they
// map and unmap memory and synchronise the device), which was ~10 % of an end-to-end step at
// config C2.  dev_free waits for the context's stream -- every kernel that touches a context
// buffer runs there -- and keeps the block; dev_alloc hands out a cached block of exactly the
// requested size when there is one.  Blocks above 1 GiB, or beyond a total of
// ICQB_ALLOC_CACHE_MB (default 2048, 0 = cache off), go back to the driver; a failed cudaMalloc
// empties the cache and retries once.
namespace {

struct BlockCache {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> blocks;   // free blocks
    std::map<void*, std::pair<int, size_t>> live;          // blocks handed out by dev_alloc_raw
    size_t cached = 0;
};

BlockCache& block_cache() {
    static BlockCache* c = new BlockCache();   // never destroyed: contexts may outlive static teardown
    return *c;
}

size_t cache_limit_bytes() {
    static const size_t limit = [] {
        const char* e = getenv("ICQB_ALLOC_CACHE_MB");
        const long long mb = e ? atoll(e) : 2048;
        return (size_t)(mb > 0 ? mb : 0) << 20;
    }();
    return limit;
}

constexpr size_t kCacheMaxBlock = (size_t)1 << 30;

// give every cached block of `device` (or of all devices: -1) back to the driver
void cache_release(int device) {
    BlockCache& C = block_cache();
    std::lock_guard<std::mutex> lock(C.m);
    int current = 0;
    cudaGetDevice(&current);
    for (auto it = C.blocks.begin(); it != C.blocks.end();) {
        if (device >= 0 && it->first.first != device) {
            ++it;
            continue;
        }
        cudaSetDevice(it->first.first);
        cudaFree(it->second);
        C.cached -= it->first.second;
        it = C.blocks.erase(it);
    }
    cudaSetDevice(current);
}

}  // namespace

cudaError_t dev_alloc_raw(icqb_ctx* ctx, void** p, size_t bytes) {
    BlockCache& C = block_cache();
    if (bytes == 0) bytes = 8;
    {
        std::lock_guard<std::mutex> lock(C.m);
        auto it = C.blocks.find({ctx->device, bytes});
        if (it != C.blocks.end()) {
            *p = it->second;
            C.cached -= bytes;
            C.blocks.erase(it);
            C.live[*p] = {ctx->device, bytes};
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();   // clear the sticky-free error state
        cache_release(-1);
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(C.m);
        C.live[*p] = {ctx->device, bytes};
    }
    return e;
}

void dev_free(icqb_ctx* ctx, void* p) {
    if (!p) return;
    BlockCache& C = block_cache();
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(C.m);
        auto it = C.live.find(p);
        if (it != C.live.end()) {
            bytes = it->second.second;
            C.live.erase(it);
        }
    }
    if (bytes == 0 || bytes > kCacheMaxBlock || cache_limit_bytes() == 0) {
        cudaFree(p);   // (synchronises the device)
        return;
    }
    // the block may still be in use by kernels of this context: they all run on its stream
    if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
    std::lock_guard<std::mutex> lock(C.m);
    if (C.cached + bytes > cache_limit_bytes()) {
        cudaFree(p);
        return;
    }
    C.blocks.insert({{ctx ? ctx->device : 0, bytes}, p});
    C.cached += bytes;
}

int ensure_work(icqb_ctx* ctx, int64_t doubles) {
    if (doubles <= ctx->work_doubles) return ICQB_OK;
    dev_free(ctx, ctx->d_work);
    ctx->d_work = nullptr;
    ctx->work_doubles = 0;
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_work, sizeof(double) * doubles));
    ctx->work_doubles = doubles;
    return ICQB_OK;
}

namespace {

// One thread per (padded) triangle.  All arithmetic that decides the position of
// a quadrature point or an area is done with explicitly rounded multiplies and
// adds in the reference's operation order (no FMA contraction), so the points
// and areas are bit-identical to the reference's doubles.
__global__ void k_prepare(const double* __restrict__ xyz, const int64_t* __restrict__ tri,
                          long long n, long long npad, long long nalloc,
                          double* __restrict__ verts, double* __restrict__ qp_soa,
                          double* __restrict__ qp_aos, double* __restrict__ area2,
                          double* __restrict__ nrm) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nalloc) return;
    if (t >= n) {
        // padding: never stored by the assembly; keep it finite and distinct
        for (int k = 0; k < kQpStride; ++k) {
            qp_aos[t * kQpStride + k] = 1.0e30;
            if (t < npad) qp_soa[(long long)k * npad + t] = -1.0e30;
        }
        area2[t] = 1.0;
        nrm[4 * t + 0] = nrm[4 * t + 1] = nrm[4 * t + 2] = nrm[4 * t + 3] = 0.0;
        return;
    }
    double pa[3], pb[3], pc[3], db[3], dc[3];
    const long long ia = tri[3 * t + 0], ib = tri[3 * t + 1], ic = tri[3 * t + 2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pa[k] = xyz[3 * ia + k];
        pb[k] = xyz[3 * ib + k];
        pc[k] = xyz[3 * ic + k];
        db[k] = __dsub_rn(pb[k], pa[k]);
        dc[k] = __dsub_rn(pc[k], pa[k]);
        verts[9 * t + k] = pa[k];
        verts[9 * t + 3 + k] = pb[k];
        verts[9 * t + 6 + k] = pc[k];
    }
    // p = pa + xi*db + eta*dc   (bem/icqQuadrature.cpp:233-234,238-239)
    for (int q = 0; q < kNq; ++q) {
        const double xi = c_triNodes[kTri8 + q].xi, eta = c_triNodes[kTri8 + q].eta;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double p =
                __dadd_rn(__dadd_rn(pa[k], __dmul_rn(xi, db[k])), __dmul_rn(eta, dc[k]));
            qp_aos[t * kQpStride + 3 * q + k] = p;
            qp_soa[(long long)(3 * q + k) * npad + t] = p;
        }
    }
    // |db x dc|   (bem/icqLaplaceMatrices.cpp:20-28, bem/icqQuadrature.cpp:163-174)
    const double p12 = __dmul_rn(db[1], dc[2]), p20 = __dmul_rn(db[2], dc[0]);
    const double p01 = __dmul_rn(db[0], dc[1]), p21 = __dmul_rn(db[2], dc[1]);
    const double p02 = __dmul_rn(db[0], dc[2]), p10 = __dmul_rn(db[1], dc[0]);
    const double cx = __dsub_rn(p12, p21), cy = __dsub_rn(p20, p02), cz = __dsub_rn(p01, p10);
    const double a2 = __dsqrt_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz)));
    area2[t] = a2;
    nrm[4 * t + 0] = cx / a2;
    nrm[4 * t + 1] = cy / a2;
    nrm[4 * t + 2] = cz / a2;
    nrm[4 * t + 3] = 0.0;
}

void free_mesh(icqb_ctx* ctx) {
    dev_free(ctx, ctx->d_xyz);
    dev_free(ctx, ctx->d_tri);
    dev_free(ctx, ctx->d_verts);
    dev_free(ctx, ctx->d_qp_soa);
    dev_free(ctx, ctx->d_qp_aos);
    dev_free(ctx, ctx->d_area2);
    dev_free(ctx, ctx->d_nrm);
    dev_free(ctx, ctx->d_bound);
    ctx->d_bound = nullptr;
    ctx->d_xyz = ctx->d_verts = ctx->d_qp_soa = ctx->d_qp_aos = ctx->d_area2 = ctx->d_nrm = nullptr;
    ctx->d_tri = nullptr;
}

}  // namespace

int launch_prepare(icqb_ctx* ctx) {
    const long long nalloc = ctx->ntri_pad + kPad;
    const int threads = 128;
    const unsigned blocks = (unsigned)((nalloc + threads - 1) / threads);
    cudaEventRecord(ctx->ev0, ctx->stream);
    k_prepare<<<blocks, threads, 0, ctx->stream>>>(ctx->d_xyz, ctx->d_tri, ctx->ntri, ctx->ntri_pad,
                                                  nalloc, ctx->d_verts, ctx->d_qp_soa,
                                                  ctx->d_qp_aos, ctx->d_area2, ctx->d_nrm);
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->launches += 1;
    ICQB_CUDA(ctx, cudaGetLastError());
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[kPrepare] = ms;
    return ICQB_OK;
}

}  // namespace icqb

using namespace icqb;

extern "C" {

int icqb_create(icqb_ctx** out, int device) {
    if (!out) return set_error(nullptr, ICQB_EARG, "icqb_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_error(nullptr, ICQB_ECUDA,
                         std::string("icqb_create: no CUDA device (") +
                             (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                             "); this library has no CPU path");
    if (device < 0 || device >= ndev)
        return set_error(nullptr, ICQB_EARG, "icqb_create: device index out of range");
    icqb_ctx* ctx = new icqb_ctx();
    ctx->device = device;
    int st = check_cuda(nullptr, cudaSetDevice(device), "cudaSetDevice");
    if (st == ICQB_OK)
        st = check_cuda(nullptr, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking),
                        "cudaStreamCreate");
    ctx->stream = ctx->own_stream;
    if (st == ICQB_OK) st = check_cuda(nullptr, cudaEventCreate(&ctx->ev0), "cudaEventCreate");
    if (st == ICQB_OK) st = check_cuda(nullptr, cudaEventCreate(&ctx->ev1), "cudaEventCreate");
    if (st == ICQB_OK)
        st = check_cuda(nullptr, cudaMallocHost(&ctx->h_pinned, 4096), "cudaMallocHost");
    if (st == ICQB_OK)
        st = check_cuda(nullptr,
                        cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device),
                        "cudaDeviceGetAttribute");
    if (st != ICQB_OK) {
        delete ctx;
        return st;
    }
    *out = ctx;
    return ICQB_OK;
}

int icqb_destroy(icqb_ctx* ctx) {
    if (!ctx) return ICQB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    symv_plan_free(ctx);
    lu_free(ctx);
    free_mesh(ctx);
    dev_free(ctx, ctx->d_ffstats);
    dev_free(ctx, ctx->d_A32);
    dev_free(ctx, ctx->d_work);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return ICQB_OK;
}

const char* icqb_last_error(const icqb_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_create_error.c_str();
}

int icqb_release_cached_memory(int device) {
    cache_release(device);
    return ICQB_OK;
}

int icqb_device(const icqb_ctx* ctx) { return ctx ? ctx->device : -1; }

uint64_t icqb_stream(const icqb_ctx* ctx) { return ctx ? (uint64_t)ctx->stream : 0; }

int icqb_set_stream(icqb_ctx* ctx, uint64_t stream) {
    if (!ctx) return ICQB_EARG;
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->own_stream;
    return ICQB_OK;
}

int icqb_synchronize(icqb_ctx* ctx) {
    if (!ctx) return ICQB_EARG;
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ICQB_OK;
}

int icqb_set_mesh(icqb_ctx* ctx, const double* xyz, int64_t npts, const int64_t* tri,
                  int64_t ntri) {
    if (!ctx) return ICQB_EARG;
    if (npts < 0 || ntri < 0 || (npts > 0 && !xyz) || (ntri > 0 && !tri))
        return set_error(ctx, ICQB_EARG, "icqb_set_mesh: bad sizes or null arrays");
    // bit-exact indexing: every vertex id must address a point
    for (int64_t k = 0; k < 3 * ntri; ++k)
        if (tri[k] < 0 || tri[k] >= npts)
            return set_error(ctx, ICQB_EARG, "icqb_set_mesh: vertex index out of range");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    free_mesh(ctx);
    symv_plan_free(ctx);
    ctx->npts = npts;
    ctx->ntri = ntri;
    ctx->ntri_pad = (ntri + kPad - 1) / kPad * kPad;
    ctx->r0 = 0;
    ctx->r1 = ntri;
    ctx->q0 = ctx->q1 = 0;
    ctx->coord_max = 0.0;
    for (int64_t k = 0; k < 3 * npts; ++k) ctx->coord_max = fmax(ctx->coord_max, fabs(xyz[k]));
    if (ntri == 0) return ICQB_OK;
    const int64_t nalloc = ctx->ntri_pad + kPad;
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_xyz, sizeof(double) * 3 * (npts > 0 ? npts : 1)));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_tri, sizeof(int64_t) * 3 * ntri));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_verts, sizeof(double) * 9 * ntri));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_qp_soa, sizeof(double) * kQpStride * ctx->ntri_pad));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_qp_aos, sizeof(double) * kQpStride * nalloc));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_area2, sizeof(double) * nalloc));
    ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_nrm, sizeof(double) * 4 * nalloc));
    ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->d_xyz, xyz, sizeof(double) * 3 * npts,
                                   cudaMemcpyHostToDevice, ctx->stream));
    ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tri, tri, sizeof(int64_t) * 3 * ntri,
                                   cudaMemcpyHostToDevice, ctx->stream));
    return launch_prepare(ctx);
}

int64_t icqb_num_triangles(const icqb_ctx* ctx) { return ctx ? ctx->ntri : 0; }

int icqb_set_rows(icqb_ctx* ctx, int64_t r0, int64_t r1) {
    if (!ctx) return ICQB_EARG;
    if (r0 < 0 || r1 < r0 || r1 > ctx->ntri)
        return set_error(ctx, ICQB_EARG, "icqb_set_rows: need 0 <= r0 <= r1 <= ntri");
    ctx->r0 = r0;
    ctx->r1 = r1;
    ctx->q0 = ctx->q1 = 0;
    return ICQB_OK;
}

int icqb_set_row_blocks(icqb_ctx* ctx, int64_t r0, int64_t r1, int64_t q0, int64_t q1) {
    if (!ctx) return ICQB_EARG;
    if (r0 < 0 || r1 < r0 || r1 > ctx->ntri || q0 < r1 || q1 < q0 || q1 > ctx->ntri)
        return set_error(ctx, ICQB_EARG,
                         "icqb_set_row_blocks: need 0 <= r0 <= r1 <= q0 <= q1 <= ntri");
    if ((r1 > r0 && (r0 % 128) != 0) || (q1 > q0 && (q0 % 128) != 0))
        return set_error(ctx, ICQB_EARG, "icqb_set_row_blocks: blocks must start on multiples of 128");
    ctx->r0 = r0;
    ctx->r1 = r1;
    ctx->q0 = q0;
    ctx->q1 = q1;
    return ICQB_OK;
}

int icqb_get_rows(const icqb_ctx* ctx, int64_t* r0, int64_t* r1) {
    if (!ctx) return ICQB_EARG;
    if (r0) *r0 = ctx->r0;
    if (r1) *r1 = ctx->r1;
    return ICQB_OK;
}

int64_t icqb_lower_num_tiles(icqb_ctx* ctx) {
    if (!ctx) return -1;
    LowerLayout L;
    if (lower_layout(ctx, &L) != ICQB_OK) return -1;
    return L.numTiles();
}

uint64_t icqb_area2_dev(const icqb_ctx* ctx) { return ctx ? (uint64_t)ctx->d_area2 : 0; }

double icqb_last_ms(const icqb_ctx* ctx, int which) {
    if (!ctx || which < 0 || which >= kNumPhases) return -1.0;
    return ctx->last_ms[which];
}

int64_t icqb_launch_count(const icqb_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
