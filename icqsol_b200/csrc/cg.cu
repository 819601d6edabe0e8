// Preconditioned conjugate gradient on a dense fp64 matrix distributed over the ranks
// of the communicator (row blocks, or the block-lower-triangular storage of the Green
// matrix).
//
// Follows solvers/icqConjugateGradient.py:49-79 (same recurrences, preconditioner
// default diag(A), absolute 2-norm stopping test) with these deliberate differences,
// all documented in DESIGN.md:
//   * optional row scaling s: the iteration runs on diag(s) A x = diag(s) b.  For
//     the Green matrix s = |db x dc| makes the operator symmetric
//     (G[j,i] = G[i,j] A_i/A_j, bem/icqLaplaceMatrices.cpp:97-98); the reference
//     class applied to the raw G does not converge;
//   * the reference recomputes the true residual with a second GEMV every
//     iteration (:74); here the recurrence residual drives the test and the true
//     residual b - A x is recomputed only when the recurrence says "converged"
//     (residual replacement if they disagree) -- one GEMV per iteration;
//   * optional block preconditioner (icqb_cg_set_preconditioner): the exact inverses
//     of the 128 x 128 diagonal blocks of diag(s) A instead of 1/diag -- consecutive
//     triangles are neighbours on the surface, and on the refined spheres this cuts the
//     iteration count 2-3x (131 instead of 306 at 9,800 triangles);
//   * dot products are one partial per CTA, added in a fixed order by the CTA that
//     finishes last (reduce_util.cuh): reproducible and identical on every rank;
//   * the loop state lives on the device (CgState) and iterations are enqueued ahead
//     of the host in batches, two batches in flight, so the GPU never waits for the
//     host between iterations.
//
// Kernels per iteration, lower storage: k_symv (forms p = w + beta p on the fly),
// k_symv_reduce (stores p, z = A p and p.z), k_cg_resid (x, r, w = M^-1 r, rho, residual
// norms, beta, stopping flag).  With several ranks one NCCL all-reduce follows
// k_symv_reduce: it sums the partial products and, in the slot behind them, the ranks'
// p.z_g.  With row-block storage p = w + beta p and p.z are kernels of their own and the
// product's row slices are all-gathered.
#include <math.h>

#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"
#include "reduce_util.cuh"

namespace icqb {

namespace {

constexpr int kT = 128;                      // rows per CTA of the vector kernels == preconditioner block
constexpr int kGroups = 8;                   // thread groups of the block preconditioner product
constexpr int kGroupRows = kT / kGroups;     // 16 rows of the inverse block per group
constexpr int kVThreads = kT * kGroups;      // 1024
static_assert(kGroups == 8, "k_cg_resid joins exactly eight groups");
constexpr long long kBlockElems = (long long)kT * kT;
constexpr int kInvThreads = 1024;
constexpr int kInvLd = kT + 1;               // padded row of the in-shared-memory inversion
constexpr int kInvSmem = (kT * kInvLd + 2 * kT) * 8;   // 134,144 B

enum ResidMode { kUpdate = 0, kInit = 1, kTrue = 2 };

// every vector is full length (n) and replicated on all ranks: the ranks execute the
// same vector kernels on identical data, only the matrix product is distributed
struct CgBuf {
    double *p, *r, *w, *z, *minv, *s;
    const double* Minv;    // [nT][128][128] inverse diagonal blocks of diag(s) A; null: diagonal
    double *part_rw, *part_res, *part_rmx, *part_a, *part_b, *part_pz;   // [nT] CTA partials
    CgState* state;
    const double* pz;      // p.(A p): &state->pz, or the slot behind z that the ranks all-reduce
    long long n;
};

// diag(A) of the locally stored rows into a zeroed full-length vector
__global__ void __launch_bounds__(256) k_cg_diag(const double* A, long long ld, long long g0,
                                                 long long g1, double* d) {
    for (long long i = g0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g1;
         i += (long long)gridDim.x * blockDim.x)
        d[i] = A[(i - g0) * ld + i];
}

// diagonal of the tile-contiguous lower storage into a zeroed full-length vector
__global__ void __launch_bounds__(256) k_cg_diag_lower(const double* A, LowerLayout L, double* d) {
    const long long rows = (L.r1 - L.r0) + (L.q1 - L.q0);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < rows;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k < (L.r1 - L.r0) ? L.r0 + k : L.q0 + (k - (L.r1 - L.r0));
        const int lt = L.stripOf(i);
        const long long I = i / kLowerTile, r = i - I * kLowerTile;
        d[i] = A[(L.tileBase(lt) + I) * kLowerTileElems + r * kLowerTile + r];
    }
}

// Diagonal 128 x 128 blocks of diag(scale) A, one CTA per block, identity-padded beyond n.
// Blocks (rows) this rank does not hold are written as zeros, so that an all-reduce
// assembles the complete set on every rank.
__global__ void __launch_bounds__(256) k_cg_blocks_lower(const double* A, LowerLayout L,
                                                         const double* scale, long long n,
                                                         double* out) {
    const long long T = blockIdx.x;
    const int lt = L.stripOf(T * kT);
    double* o = out + T * kBlockElems;
    const double* a = lt >= 0 ? A + (L.tileBase(lt) + T) * kLowerTileElems : nullptr;
    for (int e = threadIdx.x; e < kT * kT; e += blockDim.x) {
        const int r = e >> 7, c = e & (kT - 1);
        const long long i = T * kT + r, j = T * kT + c;
        double v = 0.0;
        if (a) {
            if (i < n && j < n)
                v = (scale ? scale[i] : 1.0) * a[e];
            else
                v = (r == c) ? 1.0 : 0.0;
        }
        o[e] = v;
    }
}

__global__ void __launch_bounds__(256) k_cg_blocks_full(const double* A, long long ld,
                                                        long long r0, long long r1,
                                                        const double* scale, long long n,
                                                        double* out) {
    const long long T = blockIdx.x;
    double* o = out + T * kBlockElems;
    const bool ownsLast = (n - 1 >= r0 && n - 1 < r1);   // this rank pads the last block
    for (int e = threadIdx.x; e < kT * kT; e += blockDim.x) {
        const int r = e >> 7, c = e & (kT - 1);
        const long long i = T * kT + r, j = T * kT + c;
        double v = 0.0;
        if (i < n) {
            if (i >= r0 && i < r1 && j < n) v = (scale ? scale[i] : 1.0) * A[(i - r0) * ld + j];
        } else if (ownsLast) {
            v = (r == c) ? 1.0 : 0.0;
        }
        o[e] = v;
    }
}

// In-place inverse of every 128 x 128 block: Gauss-Jordan without pivoting in shared
// memory (the blocks are principal sub-matrices of a definite matrix, every pivot has
// the sign of the matrix), then symmetrised -- a CG preconditioner must be symmetric.
__global__ void __launch_bounds__(kInvThreads) k_cg_invert_blocks(double* blocks) {
    extern __shared__ double s_inv[];
    double* a = s_inv;                    // [128][129]
    double* colk = a + kT * kInvLd;       // [128]
    double* rowk = colk + kT;             // [128]
    double* g = blocks + (long long)blockIdx.x * kBlockElems;
    const int tid = threadIdx.x;
    for (int e = tid; e < kT * kT; e += kInvThreads) a[(e >> 7) * kInvLd + (e & (kT - 1))] = g[e];
    __syncthreads();
    for (int k = 0; k < kT; ++k) {
        const double pivinv = 1.0 / a[k * kInvLd + k];
        if (tid < kT) {
            colk[tid] = a[tid * kInvLd + k];
            rowk[tid] = (tid == k) ? pivinv : a[k * kInvLd + tid] * pivinv;
        }
        __syncthreads();
        for (int e = tid; e < kT * kT; e += kInvThreads) {
            const int i = e >> 7, j = e & (kT - 1);
            double v = a[i * kInvLd + j];
            if (i == k)
                v = rowk[j];
            else if (j == k)
                v = -colk[i] * pivinv;
            else
                v = fma(-colk[i], rowk[j], v);
            a[i * kInvLd + j] = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < kT * kT; e += kInvThreads) {
        const int i = e >> 7, j = e & (kT - 1);
        g[e] = 0.5 * (a[i * kInvLd + j] + a[j * kInvLd + i]);
    }
}

// p = w + beta p  (row-block storage; the lower storage forms p inside the product)
__global__ void __launch_bounds__(256) k_cg_dir(CgBuf B) {
    if (B.state->done) return;
    const double beta = B.state->beta;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B.n;
         i += (long long)gridDim.x * blockDim.x)
        B.p[i] = fma(beta, B.p[i], B.w[i]);
}

// p.z: one partial per 128 entries, the last CTA adds them into state->pz
__global__ void __launch_bounds__(kT) k_cg_dot(CgBuf B) {
    __shared__ double s4[4];
    __shared__ double s_fin[kT];
    __shared__ int s_last;
    CgState* S = B.state;
    if (S->done) return;
    const int c = threadIdx.x;
    const long long i = (long long)blockIdx.x * kT + c;
    const double v = (i < B.n) ? B.p[i] * B.z[i] : 0.0;
    warp_part<false>(v, s4, c);
    __syncthreads();
    if (c == 0) {
        B.part_pz[blockIdx.x] = join4<false>(s4);
        s_last = take_ticket(&S->ticket_dot, gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double pz = final_reduce<false>(B.part_pz, (int)gridDim.x, s_fin, c, kT);
    if (c == 0) {
        S->pz = pz;
        S->ticket_dot = 0;
        __threadfence();
    }
}

// The residual kernel, one CTA per 128 rows (= one preconditioner block):
//   kUpdate  alpha = rho_k / (p.z); x += alpha p; r -= alpha z          (:68-72 of the reference)
//   kInit    r = s b - z (z = s A x0); p = 0; 1/(s diag) if the preconditioner is diagonal
//   kTrue    t = b - z (z = A x, unscaled): |t|_2^2, max|t|; if `replace`, r = s t
// then w = M^-1 r (block: the stored inverse block times r, four thread groups of 32 rows
// each, summed in a fixed order; diagonal: r * minv), partials of rho = r.w and of the
// norms of r/s.  The last CTA to finish closes the step: sums the partials, sets
// rho_{k+1}, beta, the residual norms, counts the iteration and raises `done`.
__global__ void __launch_bounds__(kVThreads) k_cg_resid(int mode, int k, int replace,
                                                        const double* b, const double* scale,
                                                        const double* precond, double* x, CgBuf B) {
    __shared__ double s_r[kT];
    __shared__ double s_acc[kGroups][kT];
    __shared__ double s_w4[5][4];
    __shared__ double s_fin[3 * kVThreads];
    __shared__ int s_last;
    CgState* S = B.state;
    if (mode == kUpdate && S->done) return;
    const int c = threadIdx.x, g = threadIdx.y, tid = g * kT + c;
    const long long T = blockIdx.x;
    const long long i = T * kT + c;
    const bool own = (g == 0) && (i < B.n);
    double r = 0.0, ru = 0.0, a1 = 0.0, a2 = 0.0, w = 0.0;
    const bool needW = (mode != kTrue) || replace;   // uniform
    // the inverse block is symmetric: column c of rows 16g..16g+15, coalesced over c; all
    // sixteen loads of a thread are issued before anything else so that they overlap the
    // vector update
    double m[kGroupRows];
    if (needW && B.Minv) {
        const double* M = B.Minv + T * kBlockElems + (long long)(g * kGroupRows) * kT + c;
#pragma unroll
        for (int rr = 0; rr < kGroupRows; ++rr) m[rr] = __ldg(M + rr * kT);
    }
    if (own) {
        if (mode == kUpdate) {
            const double alpha = S->rho[k & 1] / *B.pz;
            r = fma(-alpha, B.z[i], B.r[i]);
            x[i] = fma(alpha, B.p[i], x[i]);
            B.r[i] = r;
            ru = r / B.s[i];
        } else if (mode == kInit) {
            const double s = scale ? scale[i] : 1.0;
            const double bi = b[i];
            r = s * bi - B.z[i];
            B.s[i] = s;
            B.r[i] = r;
            B.p[i] = 0.0;
            if (!B.Minv) B.minv[i] = 1.0 / (s * precond[i]);
            ru = r / s;
            a1 = bi * bi;
            a2 = fabs(bi);
        } else {
            const double t = b[i] - B.z[i];
            a1 = t * t;
            a2 = fabs(t);
            if (replace) {
                r = B.s[i] * t;
                B.r[i] = r;
                ru = t;
            }
        }
    }
    if (needW) {
        if (B.Minv) {
            if (g == 0) s_r[c] = r;   // zero beyond n
            __syncthreads();
            double acc = 0.0;
#pragma unroll
            for (int rr = 0; rr < kGroupRows; ++rr) acc = fma(m[rr], s_r[g * kGroupRows + rr], acc);
            s_acc[g][c] = acc;
            __syncthreads();
            if (own)
                w = ((s_acc[0][c] + s_acc[1][c]) + (s_acc[2][c] + s_acc[3][c])) +
                    ((s_acc[4][c] + s_acc[5][c]) + (s_acc[6][c] + s_acc[7][c]));
        } else if (own) {
            w = r * B.minv[i];
        }
        if (own) B.w[i] = w;
    }
    if (g == 0) {
        warp_part<false>(r * w, s_w4[0], tid);
        warp_part<false>(ru * ru, s_w4[1], tid);
        warp_part<true>(fabs(ru), s_w4[2], tid);
        warp_part<false>(a1, s_w4[3], tid);
        warp_part<true>(a2, s_w4[4], tid);
    }
    __syncthreads();
    if (tid == 0) {
        if (needW) B.part_rw[T] = join4<false>(s_w4[0]);
        if (mode != kTrue) {
            B.part_res[T] = join4<false>(s_w4[1]);
            B.part_rmx[T] = join4<true>(s_w4[2]);
        }
        if (mode != kUpdate) {
            B.part_a[T] = join4<false>(s_w4[3]);
            B.part_b[T] = join4<true>(s_w4[4]);
        }
        s_last = take_ticket(&S->ticket_res, gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;

    // the last CTA closes the step
    __threadfence();
    const int nP = (int)gridDim.x;
    double rho = 0.0, res2 = 0.0, resmax = 0.0, fa = 0.0, fb = 0.0;
    if (mode != kTrue) {
        final_reduce3(B.part_rw, B.part_res, B.part_rmx, nP, s_fin, tid, kVThreads, &rho, &res2,
                      &resmax);
    } else if (replace) {
        rho = final_reduce<false>(B.part_rw, nP, s_fin, tid, kVThreads);
    }
    if (mode != kUpdate) {
        fa = final_reduce<false>(B.part_a, nP, s_fin, tid, kVThreads);
        fb = final_reduce<true>(B.part_b, nP, s_fin, tid, kVThreads);
    }
    if (tid == 0) {
        if (mode == kUpdate) {
            S->rho[(k + 1) & 1] = rho;
            S->beta = rho / S->rho[k & 1];
            S->res2 = res2;
            S->resmax = resmax;
            S->iters += 1;
            if ((S->use_max ? resmax : res2) <= S->thr) S->done = 1;
        } else if (mode == kInit) {
            S->rho[0] = rho;
            S->beta = 0.0;
            S->res2 = res2;
            S->resmax = resmax;
            S->bb = fa;
            S->bmax = fb;
        } else {
            S->tru2 = fa;
            S->trumax = fb;
            if (replace) {   // k = index of the next iteration
                S->rho[k & 1] = rho;
                S->beta = (k > 0) ? rho / S->rho[(k - 1) & 1] : 0.0;
            }
        }
        S->ticket_res = 0;
        __threadfence();
    }
}

__global__ void k_cg_state(CgState* s, int done, long long iters, double thr, int use_max,
                           int set_iters) {
    s->done = done;
    s->ticket_dot = 0;
    s->ticket_res = 0;
    s->thr = thr;
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

static int read_state(icqb_ctx* ctx, const CgState* dstate, CgState* out) {
    ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, dstate, sizeof(CgState), cudaMemcpyDeviceToHost,
                                   ctx->stream));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = *reinterpret_cast<const CgState*>(ctx->h_pinned);
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
    long long products = 0;    // products enqueued (iterations enqueued behind `done` included)

    // scale != null: out = diag(area or scale) A v ; null: out = A v.
    // `extra` doubles behind out[n] ride along in the all-reduce (the fused p.z).
    int apply(const double* v, double* out, const double* scale, cudaStream_t st,
              const int* skip = nullptr, const SymvFuse* fuse = nullptr, int extra = 0) {
        int rc;
        ++products;
        if (storage == 1) {
            rc = launch_symv(ctx, A, v, area, scale ? area : nullptr,
                             scale ? nullptr : inv_area, out, st, skip, fuse);
            if (rc != ICQB_OK) return rc;
            return comm_allreduce_sum(ctx, out, n + extra, st);
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

extern "C" int icqb_cg_set_preconditioner(icqb_ctx* ctx, int kind) {
    if (!ctx) return ICQB_EARG;
    if (kind != 0 && kind != 1)
        return set_error(ctx, ICQB_EARG, "icqb_cg_set_preconditioner: kind must be 0 (diagonal) or 1 (blocks)");
    ctx->precond_kind = kind;
    return ICQB_OK;
}

extern "C" int icqb_cg_last_stats(const icqb_ctx* ctx, int64_t* replacements, int64_t* products) {
    if (!ctx) return ICQB_EARG;
    if (replacements) *replacements = ctx->cg_replacements;
    if (products) *products = ctx->cg_products;
    return ICQB_OK;
}

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
    if (storage == 1 && scale && scale != ctx->d_area2)
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: with lower storage the row scaling must be icqb_area2_dev()");

    // work space
    const long long nT = (n + kT - 1) / kT;
    const bool blockPrec = (!precond) && ctx->precond_kind == 1;
    if (blockPrec && storage == 1 &&
        ((ctx->r1 % kT != 0 && ctx->r1 != n) || (ctx->q1 > ctx->q0 && ctx->q1 % kT != 0 && ctx->q1 != n)))
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: the block preconditioner needs row blocks that end on "
                         "multiples of 128 (or at n)");
    const long long nfull = chunk * P;
    const long long stateDoubles = (sizeof(CgState) + 7) / 8;
    const long long small = 8 * n + 8 + nfull + 6 * nT + stateDoubles;
    const long long need = small + (blockPrec ? nT * kBlockElems : 0) + 64;
    int rc = ensure_work(ctx, need);
    if (rc != ICQB_OK) return rc;
    CgBuf B;
    double* wp = ctx->d_work;
    B.p = wp; wp += n;
    B.r = wp; wp += n;
    B.w = wp; wp += n;
    B.z = wp; wp += n + 8;   // + the all-reduced p.z slot of the multi-rank lower storage
    B.minv = wp; wp += n;
    B.s = wp; wp += n;
    double* diagA = wp; wp += n;
    double* inv_area = wp; wp += n;
    double* gathered = wp; wp += nfull;
    B.part_rw = wp; wp += nT;
    B.part_res = wp; wp += nT;
    B.part_rmx = wp; wp += nT;
    B.part_a = wp; wp += nT;
    B.part_b = wp; wp += nT;
    B.part_pz = wp; wp += nT;
    B.state = reinterpret_cast<CgState*>(wp); wp += stateDoubles;
    double* blocks = wp;
    B.Minv = nullptr;
    B.n = n;

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

    // fusion of the direction update and of p.z into the symmetric product.  With several
    // ranks every rank forms p.(its partial product) and the sum over the ranks rides in the
    // all-reduce of the partial products, one slot behind z: p.z = sum_g p.z_g.
    SymvFuse fuse;
    const bool fuseDir = (storage == 1);
    B.pz = (fuseDir && P > 1) ? B.z + n : &B.state->pz;
    if (fuseDir) {
        fuse.w = B.w;
        fuse.beta = &B.state->beta;
        fuse.p = B.p;
        fuse.pz_part = B.part_pz;
        fuse.pz = const_cast<double*>(B.pz);
        fuse.ticket = &B.state->ticket_dot;
    }

    cudaEvent_t evTotal0 = ctx->ev0, evTotal1 = ctx->ev1;
    cudaEvent_t evg0[2] = {nullptr, nullptr}, evg1[2] = {nullptr, nullptr}, evb[2] = {nullptr, nullptr};
    for (int q = 0; q < 2; ++q) {
        ICQB_CUDA(ctx, cudaEventCreate(&evg0[q]));
        ICQB_CUDA(ctx, cudaEventCreate(&evg1[q]));
        ICQB_CUDA(ctx, cudaEventCreateWithFlags(&evb[q], cudaEventDisableTiming));
    }
    double gemv_ms = 0.0;
    long long gemv_count = 0;
    int status = ICQB_OK;
    const dim3 vblock(kT, kGroups);
    const unsigned vgrid = (unsigned)nT;
    const int* skip = reinterpret_cast<const int*>(B.state);   // &state->done
    CgState* hslot = reinterpret_cast<CgState*>(ctx->h_pinned);   // two read-back slots

    // one batch of iterations kstart .. kstart+nb-1, then the state into read-back slot `slot`
    auto enqueue_batch = [&](long long kstart, long long nb, int slot) -> int {
        for (long long i = 0; i < nb; ++i) {
            const int kk = (int)((kstart + i) & 0x3fffffff);
            if (!fuseDir) {
                k_cg_dir<<<128, 256, 0, st>>>(B);
                ctx->launches += 1;
            }
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg0[slot], st));
            int r2 = mv.apply(B.p, B.z, scale, st, skip, fuseDir ? &fuse : nullptr, fuseDir ? 1 : 0);
            if (r2 != ICQB_OK) return r2;
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg1[slot], st));
            if (!fuseDir) {
                k_cg_dot<<<vgrid, kT, 0, st>>>(B);
                ctx->launches += 1;
            }
            k_cg_resid<<<vgrid, vblock, 0, st>>>(kUpdate, kk, 0, b, scale, precond, x, B);
            ctx->launches += 1;
        }
        ICQB_CUDA(ctx, cudaGetLastError());
        ICQB_CUDA(ctx, cudaMemcpyAsync(&hslot[slot], B.state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
        ICQB_CUDA(ctx, cudaEventRecord(evb[slot], st));
        return ICQB_OK;
    };

#define CG_TRY(expr)                       \
    do {                                   \
        status = (expr);                   \
        if (status != ICQB_OK) goto done;  \
    } while (0)
#define CG_CUDA(expr) CG_TRY(check_cuda(ctx, (expr), #expr))

    {
        CG_CUDA(cudaEventRecord(evTotal0, st));
        k_zero<<<128, 256, 0, st>>>(ctx->d_work, small);
        ctx->launches += 1;
        if (storage == 1) {
            k_recip<<<128, 256, 0, st>>>(ctx->d_area2, inv_area, n);
            ctx->launches += 1;
        }
        LowerLayout L{0, 0, 0, 0, 0, 0, 0};
        if (storage == 1) CG_TRY(lower_layout(ctx, &L));
        if (blockPrec) {
            // inverse diagonal blocks of diag(s) A, replicated on every rank
            if (storage == 1)
                k_cg_blocks_lower<<<vgrid, 256, 0, st>>>(A, L, scale, n, blocks);
            else
                k_cg_blocks_full<<<vgrid, 256, 0, st>>>(A, ld, mv.r0, mv.r1, scale, n, blocks);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
            CG_TRY(comm_allreduce_sum(ctx, blocks, nT * kBlockElems, st));
            CG_CUDA(cudaFuncSetAttribute(k_cg_invert_blocks,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kInvSmem));
            k_cg_invert_blocks<<<vgrid, kInvThreads, kInvSmem, st>>>(blocks);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
            B.Minv = blocks;
        } else if (!precond) {
            // preconditioner: diag(A) unless given
            if (storage == 1) {
                k_cg_diag_lower<<<128, 256, 0, st>>>(A, L, diagA);
                ctx->launches += 1;
            } else if (mv.r1 > mv.r0) {
                k_cg_diag<<<128, 256, 0, st>>>(A, ld, mv.r0, mv.r1, diagA);
                ctx->launches += 1;
            }
            CG_TRY(comm_allreduce_sum(ctx, diagA, n, st));
            precond = diagA;
        }
        // z = s * (A x0); r, w, rho_0 and the norms of b
        CG_TRY(mv.apply(x, B.z, scale, st));
        k_cg_resid<<<vgrid, vblock, 0, st>>>(kInit, 0, 0, b, scale, precond, x, B);
        ctx->launches += 1;
        CG_CUDA(cudaGetLastError());

        CgState hs;
        CG_TRY(read_state(ctx, B.state, &hs));
        // stopping test (rel): 0 absolute 2-norm (the reference class), 1 relative 2-norm,
        // 2 relative max norm  max|b - A x| <= tol * max|b|  (the backward-error criterion)
        const bool use_max = (rel == 2);
        const double thr = use_max ? tol * hs.bmax : (rel ? tol * sqrt(hs.bb) : tol);
        const double thr_dev = use_max ? thr : thr * thr;
        double err = use_max ? hs.resmax : sqrt(hs.res2);
        long long k = 0;
        int replacements = 0;
        double true2 = sqrt(hs.res2), trueinf = hs.resmax, prev_true = -1.0;
        bool breakdown = false;
        k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 1);
        ctx->launches += 1;
        const long long batch = 8;

        while (true) {
            if (err <= thr || k >= maxit) {
                // confirm with the true residual b - A x (and stop, or replace r)
                CG_TRY(mv.apply(x, B.z, nullptr, st));
                const bool last = (k >= maxit) || replacements >= 8;
                k_cg_resid<<<vgrid, vblock, 0, st>>>(kTrue, (int)(k & 0x3fffffff), last ? 0 : 1, b,
                                                     scale, precond, x, B);
                ctx->launches += 1;
                CG_CUDA(cudaGetLastError());
                CG_TRY(read_state(ctx, B.state, &hs));
                true2 = sqrt(hs.tru2);
                trueinf = hs.trumax;
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
            // iterations are enqueued `batch` at a time, two batches in flight; the device
            // decides when to stop (CgState), the host only reads the state behind it
            long long enq = k;
            int inflight = 0, head = 0;
            bool stop = false;
            while (true) {
                while (!stop && inflight < 2 && enq < maxit) {
                    const long long nb = std::min<long long>(batch, maxit - enq);
                    CG_TRY(enqueue_batch(enq, nb, (head + inflight) & 1));
                    enq += nb;
                    ++inflight;
                }
                if (inflight == 0) break;
                CG_CUDA(cudaEventSynchronize(evb[head]));
                hs = hslot[head];
                float gms = 0;
                if (cudaEventElapsedTime(&gms, evg0[head], evg1[head]) == cudaSuccess) {
                    gemv_ms += gms;
                    ++gemv_count;
                }
                head ^= 1;
                --inflight;
                k = hs.iters;
                err = use_max ? hs.resmax : sqrt(hs.res2);
                if (!(hs.res2 == hs.res2)) breakdown = true;   // NaN
                if (hs.done || breakdown) stop = true;   // later batches are no-ops: drain them
            }
            if (breakdown) {  // singular / indefinite system
                true2 = trueinf = nan("");
                break;
            }
        }
        CG_CUDA(cudaEventRecord(evTotal1, st));
        CG_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, evTotal0, evTotal1);
        ctx->last_ms[kCg] = ms;
        if (gemv_count) ctx->last_ms[kGemv] = gemv_ms / (double)gemv_count;
        ctx->cg_replacements = replacements;
        ctx->cg_products = mv.products;
        if (iters) *iters = k;
        if (resid2) *resid2 = true2;
        if (residinf) *residinf = trueinf;
    }
done:
    if (status != ICQB_OK) cudaStreamSynchronize(st);
    for (int q = 0; q < 2; ++q) {
        cudaEventDestroy(evg0[q]);
        cudaEventDestroy(evg1[q]);
        cudaEventDestroy(evb[q]);
    }
    return status;
#undef CG_TRY
#undef CG_CUDA
}
