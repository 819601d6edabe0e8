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

struct CgBuf {
    double *pfull, *xfull, *r, *w, *z, *minv, *s;
    double *pz;        // [kNB]
    double *rw;        // [2][kNB]
    double *res;       // [kNB]  sum (r/s)^2  (recurrence residual of the unscaled system)
    double *tru;       // [kNB]  sum (b - A x)^2
    double *bb;        // [kNB]  sum b^2
    double *mx;        // [nranks] max |b - A x| per rank
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

// r = s*(b - A x0) given z = s*(A x0); minv = 1/(s*precond); w = r*minv; p = 0
__global__ void __launch_bounds__(kVT) k_cg_init(long long m, long long r0, const double* A,
                                                 long long ld, const double* b,
                                                 const double* scale, const double* precond,
                                                 CgBuf B) {
    __shared__ double s_red[kVT / 32];
    double rw = 0.0, rs = 0.0, bb = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < m; i += (long long)kNB * kVT) {
        const double s = scale ? scale[i] : 1.0;
        const double d = precond ? precond[i] : A[i * ld + r0 + i];
        const double r = s * b[i] - B.z[i];
        const double mi = 1.0 / (s * d);
        const double w = r * mi;
        B.s[i] = s;
        B.minv[i] = mi;
        B.r[i] = r;
        B.w[i] = w;
        B.pfull[r0 + i] = 0.0;
        rw = fma(r, w, rw);
        const double ru = r / s;
        rs = fma(ru, ru, rs);
        bb = fma(b[i], b[i], bb);
    }
    double t = block_sum(rw, s_red);
    if (threadIdx.x == 0) B.rw[blockIdx.x] = t;
    t = block_sum(rs, s_red);
    if (threadIdx.x == 0) B.res[blockIdx.x] = t;
    t = block_sum(bb, s_red);
    if (threadIdx.x == 0) B.bb[blockIdx.x] = t;
}

// p = w + beta p, beta = rho_k / rho_{k-1} (0 on the first iteration)
__global__ void __launch_bounds__(kVT) k_cg_dir(long long m, long long r0, int k, CgBuf B) {
    double beta = 0.0;
    if (k > 0) beta = sum_partials(B.rw + (k & 1) * kNB) / sum_partials(B.rw + ((k - 1) & 1) * kNB);
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < m; i += (long long)kNB * kVT)
        B.pfull[r0 + i] = fma(beta, B.pfull[r0 + i], B.w[i]);
}

// partial p.z over the local rows
__global__ void __launch_bounds__(kVT) k_cg_dot(long long m, long long r0, CgBuf B) {
    __shared__ double s_red[kVT / 32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < m; i += (long long)kNB * kVT)
        acc = fma(B.pfull[r0 + i], B.z[i], acc);
    const double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) B.pz[blockIdx.x] = t;
}

// alpha = rho_k / (p.z); x += alpha p; r -= alpha z; w = r/M; partial rho_{k+1}, |r/s|^2
__global__ void __launch_bounds__(kVT) k_cg_update(long long m, long long r0, int k, double* x,
                                                   CgBuf B) {
    __shared__ double s_red[kVT / 32];
    const double alpha = sum_partials(B.rw + (k & 1) * kNB) / sum_partials(B.pz);
    double rw = 0.0, rs = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < m; i += (long long)kNB * kVT) {
        const double r = fma(-alpha, B.z[i], B.r[i]);
        const double w = r * B.minv[i];
        x[i] = fma(alpha, B.pfull[r0 + i], x[i]);
        B.r[i] = r;
        B.w[i] = w;
        rw = fma(r, w, rw);
        const double ru = r / B.s[i];
        rs = fma(ru, ru, rs);
    }
    double t = block_sum(rw, s_red);
    if (threadIdx.x == 0) B.rw[((k + 1) & 1) * kNB + blockIdx.x] = t;
    t = block_sum(rs, s_red);
    if (threadIdx.x == 0) B.res[blockIdx.x] = t;
}

// true residual t = b - A x given z = A x (unscaled); optionally replace r, w, rho
__global__ void __launch_bounds__(kVT) k_cg_true(long long m, int rank, int nranks, int replace,
                                                 int knext, const double* b, CgBuf B) {
    __shared__ double s_red[kVT / 32];
    double tt = 0.0, mx = 0.0, rw = 0.0;
    for (long long i = (long long)blockIdx.x * kVT + threadIdx.x; i < m; i += (long long)kNB * kVT) {
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
    if (threadIdx.x == 0) {
        // per-rank maxima live in a [nranks] array that is sum-all-reduced:
        // a rank only ever contributes to its own slot
        if (blockIdx.x == 0)
            for (int q = 0; q < nranks; ++q)
                if (q != rank) B.mx[q] = 0.0;
        // combine blocks through an ordered integer max (values are >= 0)
        atomicMax(reinterpret_cast<unsigned long long*>(B.mx + rank),
                  (unsigned long long)__double_as_longlong(t));
    }
}

__global__ void k_zero(double* p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = 0.0;
}

}  // namespace

// full[off .. off+m) = local, then all-gather the chunk-sized slots in place
static int gather_local(icqb_ctx* ctx, const double* local, long long m, double* full,
                        long long off, long long chunk, cudaStream_t st) {
    if (m > 0)
        ICQB_CUDA(ctx, cudaMemcpyAsync(full + off, local, sizeof(double) * m,
                                       cudaMemcpyDeviceToDevice, st));
    if (comm_nranks(ctx) > 1) return comm_allgather(ctx, full + off, full, chunk, st);
    return ICQB_OK;
}

static int read_sum(icqb_ctx* ctx, const double* dpart, int count, double* out) {
    ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, dpart, sizeof(double) * count,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    for (int k = 0; k < count; ++k) t += ctx->h_pinned[k];
    *out = t;
    return ICQB_OK;
}

}  // namespace icqb

using namespace icqb;

extern "C" int icqb_cg_solve_dev(icqb_ctx* ctx, uint64_t dA_, int64_t n, int64_t ld,
                                 uint64_t drowscale_, uint64_t dprecond_, uint64_t db_,
                                 uint64_t dx_, double tol, int rel, int64_t maxit, int64_t* iters,
                                 double* resid2, double* residinf) {
    if (!ctx) return ICQB_EARG;
    if (!dA_ || !db_ || !dx_ || n <= 0 || ld < n)
        return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: null pointer or bad sizes");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double* A = reinterpret_cast<const double*>(dA_);
    const double* scale = reinterpret_cast<const double*>(drowscale_);
    const double* precond = reinterpret_cast<const double*>(dprecond_);
    const double* b = reinterpret_cast<const double*>(db_);
    double* x = reinterpret_cast<double*>(dx_);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long chunk = (n + P - 1) / P;
    const long long r0 = std::min<long long>(n, (long long)rank * chunk);
    const long long r1 = std::min<long long>(n, r0 + chunk);
    const long long m = r1 - r0;
    const long long off = (long long)rank * chunk;  // slot of this rank in the gathered vectors
    if (maxit < 0) maxit = n;
    cudaStream_t st = ctx->stream;

    // work space
    const long long nfull = chunk * P;
    const long long need = 2 * nfull + 5 * chunk + 6 * kNB + P + 64;
    int rc = ensure_work(ctx, need);
    if (rc != ICQB_OK) return rc;
    CgBuf B;
    double* wp = ctx->d_work;
    B.pfull = wp; wp += nfull;
    B.xfull = wp; wp += nfull;
    B.r = wp; wp += chunk;
    B.w = wp; wp += chunk;
    B.z = wp; wp += chunk;
    B.minv = wp; wp += chunk;
    B.s = wp; wp += chunk;
    B.pz = wp; wp += kNB;
    B.rw = wp; wp += 2 * kNB;
    B.res = wp; wp += kNB;
    B.tru = wp; wp += kNB;
    B.bb = wp; wp += kNB;
    B.mx = wp; wp += P;

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
        k_zero<<<kNB, kVT, 0, st>>>(B.pfull, 2 * nfull + 5 * chunk + 6 * kNB + P);
        ctx->launches += 1;
        // z = s * (A x0)
        CG_TRY(gather_local(ctx, x, m, B.xfull, off, chunk, st));
        CG_TRY(launch_gemv(ctx, A, m, n, ld, B.xfull, B.z, scale, st));
        k_cg_init<<<kNB, kVT, 0, st>>>(m, r0, A, ld, b, scale, precond, B);
        ctx->launches += 1;
        CG_CUDA(cudaGetLastError());
        CG_TRY(comm_allreduce_sum(ctx, B.rw, 2 * kNB + kNB, st));       // rw[0..1], res (contiguous)
        CG_TRY(comm_allreduce_sum(ctx, B.bb, kNB, st));

        double res2 = 0.0, bb = 0.0;
        CG_TRY(read_sum(ctx, B.res, kNB, &res2));
        CG_TRY(read_sum(ctx, B.bb, kNB, &bb));
        const double thr = rel ? tol * sqrt(bb) : tol;
        double err = sqrt(res2);
        long long k = 0;
        int replacements = 0;
        bool converged = false;
        double true2 = err, trueinf = 0.0;

        while (true) {
            if (err <= thr || k >= maxit) {
                // confirm with the true residual b - A x (and stop, or replace r)
                CG_TRY(gather_local(ctx, x, m, B.xfull, off, chunk, st));
                CG_TRY(launch_gemv(ctx, A, m, n, ld, B.xfull, B.z, nullptr, st));
                const bool last = (k >= maxit) || replacements >= 8;
                k_cg_true<<<kNB, kVT, 0, st>>>(m, rank, P, last ? 0 : 1, (int)(k & 1), b, B);
                ctx->launches += 1;
                CG_CUDA(cudaGetLastError());
                CG_TRY(comm_allreduce_sum(ctx, B.tru, kNB, st));
                CG_TRY(comm_allreduce_sum(ctx, B.mx, P, st));
                if (!last) CG_TRY(comm_allreduce_sum(ctx, B.rw + (k & 1) * kNB, kNB, st));
                double t2 = 0.0;
                CG_TRY(read_sum(ctx, B.tru, kNB, &t2));
                CG_CUDA(cudaMemcpyAsync(ctx->h_pinned, B.mx, sizeof(double) * P,
                                        cudaMemcpyDeviceToHost, st));
                CG_CUDA(cudaStreamSynchronize(st));
                trueinf = 0.0;
                for (int q = 0; q < P; ++q) trueinf = std::max(trueinf, ctx->h_pinned[q]);
                true2 = sqrt(t2);
                // reset the per-rank max slot for a later pass
                k_zero<<<1, 32, 0, st>>>(B.mx, P);
                ctx->launches += 1;
                if (true2 <= thr) converged = true;
                if (converged || last) break;
                ++replacements;
                err = true2;  // r was replaced by the true residual; iterate on
                if (err <= thr) { converged = true; break; }
            }
            k_cg_dir<<<kNB, kVT, 0, st>>>(m, r0, (int)k, B);
            ctx->launches += 1;
            if (P > 1) CG_TRY(comm_allgather(ctx, B.pfull + off, B.pfull, chunk, st));
            CG_CUDA(cudaEventRecord(evg0, st));
            CG_TRY(launch_gemv(ctx, A, m, n, ld, B.pfull, B.z, scale, st));
            CG_CUDA(cudaEventRecord(evg1, st));
            k_cg_dot<<<kNB, kVT, 0, st>>>(m, r0, B);
            ctx->launches += 1;
            CG_TRY(comm_allreduce_sum(ctx, B.pz, kNB, st));
            k_cg_update<<<kNB, kVT, 0, st>>>(m, r0, (int)k, x, B);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
            if (P > 1) {
                CG_TRY(comm_allreduce_sum(ctx, B.rw + ((k + 1) & 1) * kNB, kNB, st));
                CG_TRY(comm_allreduce_sum(ctx, B.res, kNB, st));
            }
            CG_TRY(read_sum(ctx, B.res, kNB, &res2));
            float gms = 0;
            cudaEventElapsedTime(&gms, evg0, evg1);
            gemv_ms += gms;
            ++gemv_count;
            err = sqrt(res2);
            ++k;
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
