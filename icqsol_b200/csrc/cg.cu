// Jacobi-preconditioned conjugate gradient on a dense row-major fp64 matrix,
// row-block sharded over the ranks of the communicator.
//
// Follows solvers/icqConjugateGradient.py:49-79 (same recurrences, same
// preconditioner default diag(A), absolute 2-norm stopping test) with these
// deliberate differences, all documented in DESIGN.md:
//   * optional row scaling s: the iteration runs on diag(s) A x = diag(s) b.  For
//     the Green matrix s = |db x dc| makes the operator symmetric
//     (G[j,i] = G[i,j] A_i/A_j, bem/icqLaplaceMatrices.cpp:97-98); the reference
//     class applied to the raw G does not converge;
//   * the reference recomputes the true residual with a second GEMV every
//     iteration (:74); here the recurrence residual drives the test and the true
//     residual b - A x is recomputed only when the recurrence says "converged"
//     (residual replacement if they disagree) -- one GEMV per iteration;
//   * dot products are block partials summed in a fixed order (reproducible),
//     all-reduced across ranks as partial arrays.
#include <math.h>

#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"

namespace icqb {

namespace {

constexpr int kNB = 128;       // blocks of every vector kernel == partials per dot product
constexpr int kVT = 256;       // threads per block

// Loop state kept on the device so that iterations can be enqueued ahead of the
// host: once the recurrence residual meets the threshold every later kernel of the
// batch returns immediately, and `iters` stops counting -- the iteration count is the
// one a per-iteration host test would give (solvers/icqConjugateGradient.py:64).
struct CgState {
    int done;
    unsigned int ticket;
    long long iters;
    double thr2;       // stopping threshold: squared 2-norm, or max norm (use_max)
    double res2;       // squared 2-norm of the recurrence residual after the last iteration
    double resmax;     // its max norm
    int use_max;       // 1: the stopping test is on the max norm
    int pad;
};

// every vector is full length (n) and replicated on all ranks: the ranks execute the
// same vector kernels on identical data, only the matrix product is distributed
struct CgBuf {
    double *p, *r, *w, *z, *minv, *s, *d;
    double *pz;        // [kNB]
    double *rw;        // [2][kNB]
    double *res;       // [kNB]  sum (r/s)^2  (recurrence residual of the unscaled system)
    double *tru;       // [kNB]  sum (b - A x)^2
    double *bb;        // [kNB]  sum b^2
    double *mx;        // [kNB]  max |b - A x| per block
    double *rmx;       // [kNB]  max |r/s| per block (recurrence residual)
    double *bmx;       // [kNB]  max |b| per block
    CgState *state;    // device-side loop state (see below)
};

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kVT / 32; ++k) t += s_red[k];
    }
    __syncthreads();
    return t;  // valid on thread 0
}

__device__ __forceinline__ double block_max(double v, double* s_red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kVT / 32; ++k) t = fmax(t, s_red[k]);
    }
    __syncthreads();
    return t;
}

// every block sums the kNB partials in the same fixed order
__device__ __forceinline__ double sum_partials(const double* part) {
    double t = 0.0;
    for (int k = 0; k < kNB; ++k) t += part[k];
    return t;
}

// diag(A) of the locally stored rows into a zeroed full-length vector
__global__ void __launch_bounds__(kVT) k_cg_diag(const double* A, long long ld, long long g0,
                                                 long long g1, long long lrow0, double* d) {
    for (long long i = g0 + (long long)blockIdx.x * kVT + threadIdx.x; i < g1;
         i += (long long)kNB * kVT)
        d[i] = A[(lrow0 + (i - g0)) * ld + i];
}

// diagonal of the tile-contiguous lower storage into a zeroed full-length vector
__global__ void __launch_bounds__(kVT) k_cg_diag_lower(const double* A, LowerLayout L, double* d) {
    const long long rows = (L.r1 - L.r0) + (L.q1 - L.q0);
    for (long long k = (long long)blockIdx.x * kVT + threadIdx.x; k < rows;
         k += (long long)kNB * kVT) {
        const long long i = k < (L.r1 - L.r0) ? L.r0 + k : L.q0 + (k - (L.r1 - L.r0));
        const int lt = L.stripOf(i);
        const long long I = i / kLowerTile, r = i - I * kLowerTile;
        d[i] = A[(L.tileBase(lt) + I) * kLowerTileElems + r * kLowerTile + r];
    }
}

// r = s*b - z (z = s*(A x0)); minv = 1/(s*precond); w = r*minv; p = 0
__global__ void __launch_bounds__(kVT) k_cg_init(long long n, const double* b,
                                                 const double* scale, const double* precond,
                                                 CgBuf B) {
    __shared__ double s_red[kVT / 32];
    double rw = 0.0, rs = 0.0, bb = 0.0, rm = 0.0, bm = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < n; i += (long long)kNB * kVT) {
        const double s = scale ? scale[i] : 1.0;
        const double r = s * b[i] - B.z[i];
        const double mi = 1.0 / (s * precond[i]);
        const double w = r * mi;
        B.s[i] = s;
        B.minv[i] = mi;
        B.r[i] = r;
        B.w[i] = w;
        B.p[i] = 0.0;
        rw = fma(r, w, rw);
        const double ru = r / s;
        rs = fma(ru, ru, rs);
        rm = fmax(rm, fabs(ru));
        bb = fma(b[i], b[i], bb);
        bm = fmax(bm, fabs(b[i]));
    }
    double t = block_sum(rw, s_red);
    if (threadIdx.x == 0) B.rw[blockIdx.x] = t;
    t = block_sum(rs, s_red);
    if (threadIdx.x == 0) B.res[blockIdx.x] = t;
    t = block_sum(bb, s_red);
    if (threadIdx.x == 0) B.bb[blockIdx.x] = t;
    t = block_max(rm, s_red);
    if (threadIdx.x == 0) B.rmx[blockIdx.x] = t;
    t = block_max(bm, s_red);
    if (threadIdx.x == 0) B.bmx[blockIdx.x] = t;
}

// p = w + beta p, beta = rho_k / rho_{k-1} (0 on the first iteration)
__global__ void __launch_bounds__(kVT) k_cg_dir(long long n, int k, CgBuf B) {
    if (B.state->done) return;
    double beta = 0.0;
    if (k > 0) beta = sum_partials(B.rw + (k & 1) * kNB) / sum_partials(B.rw + ((k - 1) & 1) * kNB);
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < n; i += (long long)kNB * kVT)
        B.p[i] = fma(beta, B.p[i], B.w[i]);
}

// partial p.z
__global__ void __launch_bounds__(kVT) k_cg_dot(long long n, CgBuf B) {
    __shared__ double s_red[kVT / 32];
    if (B.state->done) return;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < n; i += (long long)kNB * kVT)
        acc = fma(B.p[i], B.z[i], acc);
    const double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) B.pz[blockIdx.x] = t;
}

// alpha = rho_k / (p.z); x += alpha p; r -= alpha z; w = r/M; partial rho_{k+1}, |r/s|^2
__global__ void __launch_bounds__(kVT) k_cg_update(long long n, int k, double* x, CgBuf B) {
    __shared__ double s_red[kVT / 32];
    __shared__ bool s_last;
    if (B.state->done) return;
    const double alpha = sum_partials(B.rw + (k & 1) * kNB) / sum_partials(B.pz);
    double rw = 0.0, rs = 0.0, rm = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < n; i += (long long)kNB * kVT) {
        const double r = fma(-alpha, B.z[i], B.r[i]);
        const double w = r * B.minv[i];
        x[i] = fma(alpha, B.p[i], x[i]);
        B.r[i] = r;
        B.w[i] = w;
        rw = fma(r, w, rw);
        const double ru = r / B.s[i];
        rs = fma(ru, ru, rs);
        rm = fmax(rm, fabs(ru));
    }
    double t = block_sum(rw, s_red);
    if (threadIdx.x == 0) B.rw[((k + 1) & 1) * kNB + blockIdx.x] = t;
    const double tm = block_max(rm, s_red);
    t = block_sum(rs, s_red);
    if (threadIdx.x == 0) {
        B.res[blockIdx.x] = t;
        B.rmx[blockIdx.x] = tm;
        __threadfence();
        s_last = (atomicAdd(&B.state->ticket, 1u) == kNB - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        // the last block to finish closes the iteration: fixed-order sum of the partials
        __threadfence();
        const volatile double* part = B.res;
        const volatile double* pmax = B.rmx;
        double res2 = 0.0, resmax = 0.0;
        for (int q = 0; q < kNB; ++q) {
            res2 += part[q];
            resmax = fmax(resmax, pmax[q]);
        }
        B.state->res2 = res2;
        B.state->resmax = resmax;
        B.state->iters += 1;
        B.state->ticket = 0;
        if ((B.state->use_max ? resmax : res2) <= B.state->thr2) B.state->done = 1;
        __threadfence();
    }
}

// true residual t = b - A x given z = A x (unscaled); optionally replace r, w, rho
__global__ void __launch_bounds__(kVT) k_cg_true(long long n, int replace, int knext,
                                                 const double* b, CgBuf B) {
    __shared__ double s_red[kVT / 32];
    double tt = 0.0, mx = 0.0, rw = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < n; i += (long long)kNB * kVT) {
        const double t = b[i] - B.z[i];
        tt = fma(t, t, tt);
        mx = fmax(mx, fabs(t));
        if (replace) {
            const double r = B.s[i] * t;
            const double w = r * B.minv[i];
            B.r[i] = r;
            B.w[i] = w;
            rw = fma(r, w, rw);
        }
    }
    double t = block_sum(tt, s_red);
    if (threadIdx.x == 0) B.tru[blockIdx.x] = t;
    if (replace) {
        t = block_sum(rw, s_red);
        if (threadIdx.x == 0) B.rw[(knext & 1) * kNB + blockIdx.x] = t;
    }
    t = block_max(mx, s_red);
    if (threadIdx.x == 0) B.mx[blockIdx.x] = t;
}

__global__ void k_cg_state(CgState* s, int done, long long iters, double thr2, int use_max,
                           int set_iters) {
    s->done = done;
    s->ticket = 0;
    s->thr2 = thr2;
    s->use_max = use_max;
    if (set_iters) s->iters = iters;
}

__global__ void k_recip(const double* a, double* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = 1.0 / a[i];
}

__global__ void k_zero(double* p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = 0.0;
}

}  // namespace

static int read_partials(icqb_ctx* ctx, const double* dpart, int count, bool take_max,
                         double* out) {
    ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, dpart, sizeof(double) * count,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    for (int k = 0; k < count; ++k) t = take_max ? std::max(t, ctx->h_pinned[k]) : t + ctx->h_pinned[k];
    *out = t;
    return ICQB_OK;
}

// The distributed product: out = diag(scale) A v on every rank (scale may be null).
//   storage 0: rank g multiplies its row block and the slices are all-gathered;
//   storage 1: block-lower-triangular rows, symmetric product, partial vectors all-reduced.
struct MatVec {
    icqb_ctx* ctx;
    const double* A;
    long long n, ld;
    int storage;
    long long r0, r1, chunk;   // storage 0: this rank's rows and the all-gather slot size
    const double* area;        // storage 1: |db x dc| (mirror rule weights), full length
    const double* inv_area;    // storage 1: 1/area (for the unscaled product)
    double* gathered;          // storage 0: [nranks*chunk] staging for the all-gather

    // scaled != 0: out = diag(area or scale) A v ; scaled == 0: out = A v
    int apply(const double* v, double* out, const double* scale, cudaStream_t st,
              const int* skip = nullptr) {
        int rc;
        if (storage == 1) {
            rc = launch_symv(ctx, A, v, area, scale ? area : nullptr,
                             scale ? nullptr : inv_area, out, st, skip);
            if (rc != ICQB_OK) return rc;
            return comm_allreduce_sum(ctx, out, n, st);
        }
        const int P = comm_nranks(ctx);
        double* dst = (P == 1) ? out : gathered + (long long)comm_rank(ctx) * chunk;
        rc = launch_gemv(ctx, A, r1 - r0, n, ld, v, dst, scale ? scale + r0 : nullptr, st, skip);
        if (rc != ICQB_OK || P == 1) return rc;
        rc = comm_allgather(ctx, dst, gathered, chunk, st);
        if (rc != ICQB_OK) return rc;
        return check_cuda(ctx, cudaMemcpyAsync(out, gathered, sizeof(double) * n,
                                               cudaMemcpyDeviceToDevice, st), "gather copy");
    }
};

}  // namespace icqb

using namespace icqb;

extern "C" int icqb_cg_solve_dev(icqb_ctx* ctx, uint64_t dA_, int64_t n, int64_t ld, int storage,
                                 uint64_t drowscale_, uint64_t dprecond_, uint64_t db_,
                                 uint64_t dx_, double tol, int rel, int64_t maxit, int64_t* iters,
                                 double* resid2, double* residinf) {
    if (!ctx) return ICQB_EARG;
    if (!dA_ || !db_ || !dx_ || n <= 0 || (storage == 0 && ld < n))
        return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: null pointer or bad sizes");
    if (storage != 0 && storage != 1)
        return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: storage must be 0 (row blocks) or 1 (lower)");
    if (storage == 1 && (ctx->ntri != n || !ctx->d_area2))
        return set_error(ctx, ICQB_ESTATE,
                         "icqb_cg_solve_dev: lower storage needs the mesh of this matrix in the context");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double* A = reinterpret_cast<const double*>(dA_);
    const double* scale = reinterpret_cast<const double*>(drowscale_);
    const double* precond = reinterpret_cast<const double*>(dprecond_);
    const double* b = reinterpret_cast<const double*>(db_);
    double* x = reinterpret_cast<double*>(dx_);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long chunk = (n + P - 1) / P;
    if (maxit < 0) maxit = n;
    cudaStream_t st = ctx->stream;

    // work space
    const long long nfull = chunk * P;
    const long long need = 8 * n + nfull + 9 * kNB + 8 + 64;
    int rc = ensure_work(ctx, need);
    if (rc != ICQB_OK) return rc;
    CgBuf B;
    double* wp = ctx->d_work;
    B.p = wp; wp += n;
    B.r = wp; wp += n;
    B.w = wp; wp += n;
    B.z = wp; wp += n;
    B.minv = wp; wp += n;
    B.s = wp; wp += n;
    B.d = wp; wp += n;
    double* inv_area = wp; wp += n;
    double* gathered = wp; wp += nfull;
    B.pz = wp; wp += kNB;
    B.rw = wp; wp += 2 * kNB;
    B.res = wp; wp += kNB;
    B.tru = wp; wp += kNB;
    B.bb = wp; wp += kNB;
    B.mx = wp; wp += kNB;
    B.rmx = wp; wp += kNB;
    B.bmx = wp; wp += kNB;
    B.state = reinterpret_cast<CgState*>(wp); wp += 8;

    MatVec mv;
    mv.ctx = ctx;
    mv.A = A;
    mv.n = n;
    mv.ld = ld;
    mv.storage = storage;
    mv.chunk = chunk;
    mv.r0 = std::min<long long>(n, (long long)rank * chunk);
    mv.r1 = std::min<long long>(n, mv.r0 + chunk);
    mv.area = ctx->d_area2;
    mv.inv_area = inv_area;
    mv.gathered = gathered;
    if (storage == 1 && scale && scale != ctx->d_area2)
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: with lower storage the row scaling must be icqb_area2_dev()");

    cudaEvent_t evTotal0 = ctx->ev0, evTotal1 = ctx->ev1;
    cudaEvent_t evg0, evg1;
    ICQB_CUDA(ctx, cudaEventCreate(&evg0));
    ICQB_CUDA(ctx, cudaEventCreate(&evg1));
    double gemv_ms = 0.0;
    long long gemv_count = 0;
    int status = ICQB_OK;
#define CG_TRY(expr)                       \
    do {                                   \
        status = (expr);                   \
        if (status != ICQB_OK) goto done;  \
    } while (0)
#define CG_CUDA(expr) CG_TRY(check_cuda(ctx, (expr), #expr))

    {
        CG_CUDA(cudaEventRecord(evTotal0, st));
        k_zero<<<kNB, kVT, 0, st>>>(ctx->d_work, need);
        ctx->launches += 1;
        if (storage == 1) {
            k_recip<<<kNB, kVT, 0, st>>>(ctx->d_area2, inv_area, n);
            ctx->launches += 1;
        }
        // preconditioner: diag(A) unless given
        if (!precond) {
            if (storage == 1) {
                LowerLayout L;
                CG_TRY(lower_layout(ctx, &L));
                k_cg_diag_lower<<<kNB, kVT, 0, st>>>(A, L, B.d);
            } else if (mv.r1 > mv.r0) {
                k_cg_diag<<<kNB, kVT, 0, st>>>(A, ld, mv.r0, mv.r1, 0, B.d);
            }
            ctx->launches += 2;
            CG_TRY(comm_allreduce_sum(ctx, B.d, n, st));
            precond = B.d;
        }
        // z = s * (A x0)
        CG_TRY(mv.apply(x, B.z, scale, st));
        k_cg_init<<<kNB, kVT, 0, st>>>(n, b, scale, precond, B);
        ctx->launches += 1;
        CG_CUDA(cudaGetLastError());

        double res2 = 0.0, bb = 0.0;
        CG_TRY(read_partials(ctx, B.res, kNB, false, &res2));
        CG_TRY(read_partials(ctx, B.bb, kNB, false, &bb));
        // stopping test (rel): 0 absolute 2-norm (the reference class), 1 relative 2-norm,
        // 2 relative max norm  max|b - A x| <= tol * max|b|  (the backward-error criterion)
        const bool use_max = (rel == 2);
        double bmax = 0.0, rmax0 = 0.0;
        CG_TRY(read_partials(ctx, B.bmx, kNB, true, &bmax));
        CG_TRY(read_partials(ctx, B.rmx, kNB, true, &rmax0));
        const double thr = use_max ? tol * bmax : (rel ? tol * sqrt(bb) : tol);
        const double thr_dev = use_max ? thr : thr * thr;
        double err = use_max ? rmax0 : sqrt(res2);
        long long k = 0;
        int replacements = 0;
        double true2 = sqrt(res2), trueinf = rmax0, prev_true = -1.0;
        const int* skip = reinterpret_cast<const int*>(B.state);   // &state->done
        k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 1);
        ctx->launches += 1;
        // iterations are enqueued `batch` at a time; the device decides when to stop
        // (CgState), the host only polls -- no host round trip per iteration
        const long long batch = 8;

        while (true) {
            if (err <= thr || k >= maxit) {
                // confirm with the true residual b - A x (and stop, or replace r)
                CG_TRY(mv.apply(x, B.z, nullptr, st));
                const bool last = (k >= maxit) || replacements >= 8;
                k_cg_true<<<kNB, kVT, 0, st>>>(n, last ? 0 : 1, (int)(k & 1), b, B);
                ctx->launches += 1;
                CG_CUDA(cudaGetLastError());
                double t2 = 0.0;
                CG_TRY(read_partials(ctx, B.tru, kNB, false, &t2));
                CG_TRY(read_partials(ctx, B.mx, kNB, true, &trueinf));
                true2 = sqrt(t2);
                const double truem = use_max ? trueinf : true2;
                if (truem <= thr || last) break;
                // attainable accuracy reached: the true residual no longer follows the
                // recurrence (it did not even halve since the last replacement)
                if (prev_true >= 0.0 && truem > 0.5 * prev_true) break;
                prev_true = truem;
                ++replacements;
                err = truem;  // r was replaced by the true residual; iterate on
                k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 0);
                ctx->launches += 1;
            }
            const long long nb = std::min<long long>(batch, maxit - k);
            for (long long i = 0; i < nb; ++i) {
                const int kk = (int)((k + i) & 0x3fffffff);
                k_cg_dir<<<kNB, kVT, 0, st>>>(n, kk, B);
                if (i == 0) CG_CUDA(cudaEventRecord(evg0, st));
                CG_TRY(mv.apply(B.p, B.z, scale ? scale : nullptr, st, skip));
                if (i == 0) CG_CUDA(cudaEventRecord(evg1, st));
                k_cg_dot<<<kNB, kVT, 0, st>>>(n, B);
                k_cg_update<<<kNB, kVT, 0, st>>>(n, kk, x, B);
                ctx->launches += 3;
            }
            CG_CUDA(cudaGetLastError());
            CG_CUDA(cudaMemcpyAsync(ctx->h_pinned, B.state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
            CG_CUDA(cudaStreamSynchronize(st));
            const CgState hs = *reinterpret_cast<const CgState*>(ctx->h_pinned);
            float gms = 0;
            cudaEventElapsedTime(&gms, evg0, evg1);
            gemv_ms += gms;
            ++gemv_count;
            k = hs.iters;
            err = use_max ? hs.resmax : sqrt(hs.res2);
            if (!(err == err)) {  // NaN: breakdown (singular / indefinite system)
                true2 = err;
                break;
            }
        }
        CG_CUDA(cudaEventRecord(evTotal1, st));
        CG_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, evTotal0, evTotal1);
        ctx->last_ms[kCg] = ms;
        if (gemv_count) ctx->last_ms[kGemv] = gemv_ms / (double)gemv_count;
        if (iters) *iters = k;
        if (resid2) *resid2 = true2;
        if (residinf) *residinf = trueinf;
    }
done:
    cudaEventDestroy(evg0);
    cudaEventDestroy(evg1);
    return status;
#undef CG_TRY
#undef CG_CUDA
}
