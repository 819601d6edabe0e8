// Kernels and host helpers of the conjugate gradient, shared by cg.cu (the solver behind
// icqb_cg_solve_dev) and cg_caps.cu (the solver of preconditioner kind 2, with larger dense blocks in the
// preconditioner).  Everything here has internal linkage: each translation unit that
// includes this header gets its own copy.  See cg.cu for the algorithm.
#pragma once

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
constexpr int kInvThreads = 512;
constexpr int kInvLd = kT + 1;               // padded row of the in-shared-memory inversion
constexpr int kInvSmem = (kT * kInvLd + 5 * kT) * 8;   // 137,216 B: staging tile + published vectors

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

// Gauss-Jordan inverse (no pivoting) of a 128 x 128 tile held in REGISTERS: 512 threads, thread
// (ty = warp, tx = lane) owns the 8 x 4 sub-block rows 8 ty.., columns 4 tx...  Per pivot k the
// warp that owns row k publishes the scaled pivot row (pivinv at position k), one lane per warp
// publishes its eight entries of column k, and after ONE barrier every thread does its 32 DFMAs
// from registers (the two published vectors are double-buffered by the parity of k, so the
// next pivot's writers cannot overtake this pivot's readers).  The position of row / column k
// inside the owners' sub-blocks is a compile-time constant (eight instantiations of the pivot
// step), so no register is ever selected at run time.  The first version kept the tile in shared
// memory and moved 4 shared-memory words per DFMA: 2.8 us per pivot, 354 us per tile (ncu,
// round 2); this one is bound by the barrier and the row warp's division.
// s_vec: [2][2][128] doubles (row / column, two parities) + [128] original diagonal.
// Returns the number of weak pivots seen by this thread (lane 0 of the row warps only).
constexpr int kGjRows = 8;
constexpr int kGjThreads = (kT / kGjRows) * 32;   // 512
static_assert(kInvThreads == kGjThreads, "the tile inversion kernels are launched with kInvThreads");

template <int J>   // k = 8 kw + J
__device__ __forceinline__ void tile_gj_pivot(double (&a)[kGjRows][4], double* s_vec, int kw, int tx, int ty,
                                              int& nweak) {
    constexpr int KC = J & 3;                 // column k inside its owner's four columns
    const int k = kGjRows * kw + J;
    const int lk = 2 * kw + (J >> 2);         // the lane that owns column k
    const double* s_diag = s_vec + 4 * kT;
    double* s_row = s_vec + (J & 1) * 2 * kT;   // parity of k
    double* s_col = s_row + kT;
    if (ty == kw) {   // the warp that owns row k
        const double piv = __shfl_sync(0xffffffffu, a[J][KC], lk);
        // a pivot that is not finite, or has lost 13 digits against the diagonal entry it
        // started from, is reported (the caller decides what to do with the tile)
        if (tx == 0 && (!(fabs(piv) >= 1e-13 * fabs(s_diag[k])) || !(fabs(piv) <= 1.0e300))) ++nweak;
        const double pivinv = 1.0 / piv;
        double o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = a[J][c] * pivinv;
        if (tx == lk) o[KC] = pivinv;
        double2* dst = reinterpret_cast<double2*>(s_row + 4 * tx);
        dst[0] = make_double2(o[0], o[1]);
        dst[1] = make_double2(o[2], o[3]);
    }
    if (tx == lk) {   // one lane per warp owns eight entries of column k
        double2* dst = reinterpret_cast<double2*>(s_col + kGjRows * ty);
#pragma unroll
        for (int r = 0; r < kGjRows; r += 2) dst[r >> 1] = make_double2(a[r][KC], a[r + 1][KC]);
    }
    __syncthreads();
    double rs[4], cs[kGjRows];
    {
        const double2 r01 = *reinterpret_cast<const double2*>(s_row + 4 * tx);
        const double2 r23 = *reinterpret_cast<const double2*>(s_row + 4 * tx + 2);
        rs[0] = r01.x; rs[1] = r01.y; rs[2] = r23.x; rs[3] = r23.y;
    }
#pragma unroll
    for (int r = 0; r < kGjRows; r += 2) {
        const double2 c2 = *reinterpret_cast<const double2*>(s_col + kGjRows * ty + r);
        cs[r] = c2.x;
        cs[r + 1] = c2.y;
    }
#pragma unroll
    for (int r = 0; r < kGjRows; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[r][c] = fma(-cs[r], rs[c], a[r][c]);
    if (tx == lk) {   // column k: -a_ik / pivot  (rs[KC] of this lane is 1 / pivot)
#pragma unroll
        for (int r = 0; r < kGjRows; ++r) a[r][KC] = -cs[r] * rs[KC];
    }
    if (ty == kw) {   // row k: the scaled pivot row, 1 / pivot on the diagonal
#pragma unroll
        for (int c = 0; c < 4; ++c) a[J][c] = rs[c];
    }
}

__device__ __forceinline__ int tile_gauss_jordan(double (&a)[kGjRows][4], double* s_vec) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double* s_diag = s_vec + 4 * kT;
    // the diagonal entries of rows 8 ty .. 8 ty + 7 sit in lanes 2 ty (rows 0-3) and 2 ty + 1
    if (tx == 2 * ty) {
#pragma unroll
        for (int r = 0; r < 4; ++r) s_diag[kGjRows * ty + r] = a[r][r];
    }
    if (tx == 2 * ty + 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) s_diag[kGjRows * ty + 4 + r] = a[4 + r][r];
    }
    __syncthreads();
    int nweak = 0;
#pragma unroll 1
    for (int kw = 0; kw < kT / kGjRows; ++kw) {
        tile_gj_pivot<0>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<1>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<2>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<3>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<4>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<5>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<6>(a, s_vec, kw, tx, ty, nweak);
        tile_gj_pivot<7>(a, s_vec, kw, tx, ty, nweak);
    }
    return nweak;
}

// tile (row pitch ld) <-> the threads' 8 x 4 sub-blocks: a warp moves one row of 1 KB at a time
__device__ __forceinline__ void tile_to_regs(const double* g, long long ld, double (&a)[kGjRows][4]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < kGjRows; ++r) {
        const double2* src = reinterpret_cast<const double2*>(g + (long long)(kGjRows * ty + r) * ld + 4 * tx);
        const double2 lo = src[0], hi = src[1];
        a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = hi.x; a[r][3] = hi.y;
    }
}

// symmetrised store: out = (a + a^T) / 2 through a [128][129] staging tile in shared memory
__device__ __forceinline__ void regs_to_tile_symmetric(const double (&a)[kGjRows][4], double* s_t, double* g,
                                                       long long ld) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < kGjRows; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s_t[(kGjRows * ty + r) * kInvLd + 4 * tx + c] = a[r][c];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kGjRows; ++r) {
        double o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = 0.5 * (a[r][c] + s_t[(4 * tx + c) * kInvLd + kGjRows * ty + r]);
        double2* dst = reinterpret_cast<double2*>(g + (long long)(kGjRows * ty + r) * ld + 4 * tx);
        dst[0] = make_double2(o[0], o[1]);
        dst[1] = make_double2(o[2], o[3]);
    }
}

// In-place inverse of every 128 x 128 block (the blocks are principal sub-matrices of a matrix
// that is definite wherever the reference's quadrature is accurate enough), then symmetrised -- a
// CG preconditioner must be symmetric.  A weak pivot (sliver triangles: DESIGN.md section 4.6)
// would give a huge, indefinite "inverse": such a block falls back to 1/diag and is counted in
// *bad (icqb_cg_last_blocks).  `tiles` (optional): only blocks with tiles[block] < 0 are single
// tiles of the partition; the others belong to dense blocks and are skipped.
__global__ void __launch_bounds__(kInvThreads) k_cg_invert_blocks(double* blocks, int* bad,
                                                                  const int* tiles = nullptr) {
    extern __shared__ double s_inv[];
    if (tiles && tiles[blockIdx.x] >= 0) return;
    double* s_t = s_inv;                     // [128][129] staging of the symmetrised store
    double* s_vec = s_inv + kT * kInvLd;     // published rows / columns, original diagonal
    __shared__ int s_broken;
    double* g = blocks + (long long)blockIdx.x * kBlockElems;
    if (threadIdx.x == 0) s_broken = 0;
    double a[kGjRows][4];
    tile_to_regs(g, kT, a);
    const int nweak = tile_gauss_jordan(a, s_vec);
    if (nweak) s_broken = 1;
    __syncthreads();
    if (s_broken) {
        const double* s_diag = s_vec + 4 * kT;
        for (int e = threadIdx.x; e < kT * kT; e += kInvThreads) {
            const int i = e >> 7, j = e & (kT - 1);
            g[e] = (i == j) ? 1.0 / s_diag[i] : 0.0;
        }
        if (threadIdx.x == 0 && bad) atomicAdd(bad, 1);
        return;
    }
    regs_to_tile_symmetric(a, s_t, g, kT);
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
