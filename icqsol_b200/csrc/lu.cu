// Dense LU with partial pivoting on the GPU(s): the reference's own solve,
// numpy.linalg.solve -> LAPACK dgesv (bem/icqLaplaceSolver.py:36), behind the conjugate
// gradient as the solver that cannot "not converge" (BaseLaplaceSolver.solveGreenSystemDirect).
//
// What is factored is S = diag(s) G with s = |db x dc|: S is symmetric
// (G[j,i] = G[i,j] A_i/A_j, bem/icqLaplaceMatrices.cpp:97-98), so the ROWS of G that the
// assembly kernels produce (row-major, contiguous) are, once scaled, the COLUMNS of S, and
//     G x = b   <=>   S x = diag(s) b.
// The host layer checks the result against the resident matrix with the product kernels and
// refines it (icqBaseLaplaceSolver.py), so the scaling costs no accuracy.
//
// Layout: 1-D block-cyclic COLUMN distribution over the ranks, blocks of 128 columns, global
// column block kb on rank kb % P as local block kb / P; a column is contiguous (pitch ld).
// Right-looking blocked LU, row pivoting (LAPACK's dgetrf order of operations):
//   per panel kb (owner only):   128 x { arg-max of the column, swap, scale, rank-1 update of
//                                the rest of the panel } -- one kernel per column; the arg-max
//                                of the NEXT column rides in the update kernel and the CTA that
//                                finishes last (ticket) closes it, no host round trip;
//                                then inv(L11) (unit lower, one CTA) and the panel is packed
//                                [pivots | inv(L11) | L21] and broadcast (NCCL) to the ranks;
//   every rank, own columns right of the panel:  row swaps, U12 = inv(L11) A12 and
//                                A22 -= L21 U12 -- both the same 128 x 64 x 128 FP64 tile kernel
//                                (DFMA pipe; B200 has no faster FP64 tensor path).
// Row swaps are applied to the columns right of the panel only; the forward substitution
// replays the panels in order (swap, solve with L11, update with L21), which is consistent
// with L21 as it was stored at the time.  Solve: forward and backward sweeps over the column
// blocks, the vector travels by broadcast from the block's owner.
//
// Algorithmic work: 2/3 n^3 flop over all ranks; memory 8 n^2 / P bytes per rank for the
// factors plus the panel buffer (8 * 128 * n).
#include <math.h>

#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"
#include "reduce_util.cuh"

namespace icqb {

namespace {

constexpr int kNb = 128;                 // panel width = column block of the distribution
constexpr int kMaxPart = 4096;           // arg-max partials (CTAs of a panel column kernel)

// device-side state of one panel factorisation
struct PanelArgs {
    double* W;             // local columns, column-major, pitch ld
    long long ld;
    long long lc0;         // first local column of the panel
    long long k0;          // first global row (= global column) of the panel
    int jb;                // columns of the panel
    long long n;
    double* rowbuf;        // [2][2][kNb]: per parity, copy of the pivot row and of row k0+j
    double* part_val;      // [kMaxPart] arg-max partials: |value|
    long long* part_idx;   // [kMaxPart]                    row
    unsigned int* ticket;
    long long* piv;        // [kNb] pivot row (global) of every panel column
    double* pivval;        // [kNb] pivot value
    int* info;             // first zero pivot (1-based), 0 = none
};

__device__ __forceinline__ bool better(double va, long long ia, double vb, long long ib) {
    // LAPACK idamax: the first of the largest
    return va > vb || (va == vb && ia < ib);
}

// The CTA that finishes last reduces the arg-max partials of panel column jn, records the
// pivot and copies the pivot row and row k0+jn of the panel (all jb columns) into
// rowbuf[jn & 1]: the update kernel of column jn reads them from there, so that no CTA reads a
// row that another one is rewriting.
__device__ void close_column(const PanelArgs& A, int jn, int nparts, double* s_val,
                             long long* s_idx, int tid, int nthreads) {
    double bv = -1.0;
    long long bi = 0x7fffffffffffffffLL;
    const volatile double* pv = A.part_val;
    const volatile long long* pi = A.part_idx;
    for (int q = tid; q < nparts; q += nthreads) {
        const double v = pv[q];
        const long long i = pi[q];
        if (better(v, i, bv, bi)) { bv = v; bi = i; }
    }
    s_val[tid] = bv;
    s_idx[tid] = bi;
    __syncthreads();
    for (int off = nthreads >> 1; off > 0; off >>= 1) {
        if (tid < off && better(s_val[tid + off], s_idx[tid + off], s_val[tid], s_idx[tid])) {
            s_val[tid] = s_val[tid + off];
            s_idx[tid] = s_idx[tid + off];
        }
        __syncthreads();
    }
    const long long p = s_idx[0];
    const long long rj = A.k0 + jn;
    double* rowP = A.rowbuf + (jn & 1) * 2 * kNb;
    double* rowJ = rowP + kNb;
    for (int c = tid; c < A.jb; c += nthreads) {
        const double* col = A.W + (A.lc0 + c) * A.ld;
        rowP[c] = col[p];
        rowJ[c] = col[rj];
    }
    if (tid == 0) {
        A.piv[jn] = p;
        const double pvv = A.W[(A.lc0 + jn) * A.ld + p];
        A.pivval[jn] = pvv;
        if (pvv == 0.0 && *A.info == 0) *A.info = (int)(rj + 1);
        *A.ticket = 0;
        __threadfence();
    }
}

// arg-max of the first panel column over rows [k0, n)
__global__ void __launch_bounds__(256) k_lu_pivot0(const PanelArgs A) {
    __shared__ double s_val[256];
    __shared__ long long s_idx[256];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const double* col = A.W + A.lc0 * A.ld;
    double bv = -1.0;
    long long bi = 0x7fffffffffffffffLL;
    for (long long r = A.k0 + (long long)blockIdx.x * 256 + tid; r < A.n; r += (long long)gridDim.x * 256) {
        const double v = fabs(col[r]);
        if (better(v, r, bv, bi)) { bv = v; bi = r; }
    }
    s_val[tid] = bv;
    s_idx[tid] = bi;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off && better(s_val[tid + off], s_idx[tid + off], s_val[tid], s_idx[tid])) {
            s_val[tid] = s_val[tid + off];
            s_idx[tid] = s_idx[tid + off];
        }
        __syncthreads();
    }
    if (tid == 0) {
        A.part_val[blockIdx.x] = s_val[0];
        A.part_idx[blockIdx.x] = s_idx[0];
        s_last = take_ticket(A.ticket, gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    close_column(A, 0, (int)gridDim.x, s_val, s_idx, tid, 256);
}

// Panel column j: rows r in (k0+j, n).  With the pivot row P (old row p) and old row J = k0+j:
//   row k0+j <- P (all panel columns; CTA 0);   row p <- J, then eliminated like the others:
//   l = a(r,j) / pivot, a(r,j) = l, a(r,c) -= l P(c) for c > j.
// Threads (128 rows, 4 column phases); the threads of phase 0 also own column j+1 after the
// update and feed the arg-max of the next column.
__global__ void __launch_bounds__(512) k_lu_col(const PanelArgs A, int j) {
    __shared__ double s_P[kNb], s_J[kNb];
    __shared__ double s_l[128];
    __shared__ double s_val[512];
    __shared__ long long s_idx[512];
    __shared__ int s_last;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 128 + tx;
    const double* rowP = A.rowbuf + (j & 1) * 2 * kNb;
    if (tid < A.jb) {
        s_P[tid] = rowP[tid];
        s_J[tid] = rowP[kNb + tid];
    }
    __syncthreads();
    const long long p = A.piv[j];
    const double pivot = A.pivval[j];
    const double pinv = (pivot != 0.0) ? 1.0 / pivot : 0.0;
    const long long rj = A.k0 + j;
    const long long r = rj + 1 + (long long)blockIdx.x * 128 + tx;
    const bool live = r < A.n;
    double* base = A.W + A.lc0 * A.ld;
    if (blockIdx.x == 0 && tid < A.jb) base[(long long)tid * A.ld + rj] = s_P[tid];
    const bool isP = live && (r == p);
    if (ty == 0) {
        double aj = 0.0;
        if (live) aj = isP ? s_J[j] : base[(long long)j * A.ld + r];
        // (dgetf2 multiplies by the reciprocal of the pivot)
        const double l = aj * pinv;
        s_l[tx] = l;
        if (live) base[(long long)j * A.ld + r] = l;
    }
    __syncthreads();
    const double l = s_l[tx];
    double next = 0.0;
    if (live) {
        if (isP)   // the left part of the swapped row
            for (int c = ty; c < j; c += 4) base[(long long)c * A.ld + r] = s_J[c];
        // all loads of the row first (independent: up to 32 in flight), then the updates
        double old[kNb / 4];
#pragma unroll
        for (int q = 0; q < kNb / 4; ++q) {
            const int c = j + 1 + ty + 4 * q;
            old[q] = 0.0;
            if (c < A.jb) old[q] = isP ? s_J[c] : base[(long long)c * A.ld + r];
        }
#pragma unroll
        for (int q = 0; q < kNb / 4; ++q) {
            const int c = j + 1 + ty + 4 * q;
            if (c < A.jb) {
                const double v = fma(-l, s_P[c], old[q]);
                base[(long long)c * A.ld + r] = v;
                if (q == 0 && ty == 0) next = v;
            }
        }
    }
    if (j + 1 >= A.jb) return;   // last column of the panel: nothing to search
    // arg-max of column j+1 over this CTA's rows (phase 0 holds the new values)
    s_val[tid] = (ty == 0 && live) ? fabs(next) : -1.0;
    s_idx[tid] = (ty == 0 && live) ? r : 0x7fffffffffffffffLL;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (tid < off && better(s_val[tid + off], s_idx[tid + off], s_val[tid], s_idx[tid])) {
            s_val[tid] = s_val[tid + off];
            s_idx[tid] = s_idx[tid + off];
        }
        __syncthreads();
    }
    if (tid == 0) {
        A.part_val[blockIdx.x] = s_val[0];
        A.part_idx[blockIdx.x] = s_idx[0];
        s_last = take_ticket(A.ticket, gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    close_column(A, j + 1, (int)gridDim.x, s_val, s_idx, tid, 512);
}

// Packed panel: [0, 128) pivot rows (as doubles), [128, 128 + 128*128) inv(L11) column-major,
// then L21: rows k0+jb .. n of the jb panel columns, column-major with pitch lp.
constexpr long long kPackInv = kNb;
constexpr long long kPackL21 = kNb + (long long)kNb * kNb;

__global__ void __launch_bounds__(256) k_lu_pack(const PanelArgs A, double* buf, long long lp) {
    const long long m2 = A.n - A.k0 - A.jb;
    if (blockIdx.x == 0 && threadIdx.x < kNb)
        buf[threadIdx.x] = threadIdx.x < A.jb ? (double)A.piv[threadIdx.x] : -1.0;
    const long long total = m2 * A.jb;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long c = e / m2, r = e - c * m2;
        buf[kPackL21 + c * lp + r] = A.W[(A.lc0 + c) * A.ld + A.k0 + A.jb + r];
    }
}

// inv(L11), L11 unit lower triangular (jb x jb, below the diagonal of the panel's first rows):
// thread j forms column j of the inverse by forward substitution, x_j = 1,
// x_i = -sum_{k=j}^{i-1} L(i,k) x_k.  L sits packed in shared memory (strictly lower part), the
// thread's column in a private row of shared memory; all threads walk (i, k) in lockstep so that
// the L reads are broadcasts.  Identity beyond jb.
constexpr int kInvLdX = kNb + 1;
constexpr int kTrinvSmem = (kNb * (kNb - 1) / 2 + kNb * kInvLdX) * 8;

__global__ void __launch_bounds__(kNb) k_lu_trinv(const PanelArgs A, double* buf) {
    extern __shared__ double s_mem[];
    double* sL = s_mem;                          // packed: (i, k), k < i  at  i (i-1)/2 + k
    double* sX = s_mem + kNb * (kNb - 1) / 2;    // [j][i], pitch 129
    const int j = threadIdx.x;
    const int jb = A.jb;
    for (int k = 0; k < jb; ++k) {
        // column k of L11, rows k+1 .. jb-1: coalesced over the threads
        if (j > k && j < jb) sL[j * (j - 1) / 2 + k] = A.W[(A.lc0 + k) * A.ld + A.k0 + j];
    }
    for (int i = 0; i < kNb; ++i) sX[j * kInvLdX + i] = (i == j) ? 1.0 : 0.0;
    __syncthreads();
    if (j < jb) {
        double* x = sX + j * kInvLdX;
        for (int i = 1; i < jb; ++i) {
            double acc = 0.0;
            const double* Li = sL + i * (i - 1) / 2;
            for (int k = 0; k < i; ++k) acc = fma(Li[k], (k >= j) ? x[k] : 0.0, acc);
            if (i > j) x[i] = -acc;
        }
    }
    __syncthreads();
    // column-major out: element (i, c) at buf[kPackInv + c * 128 + i]
    for (int c = 0; c < kNb; ++c) buf[kPackInv + (long long)c * kNb + j] = sX[c * kInvLdX + j];
}

// row swaps of one panel on local columns [c0, c1): one thread per column, the swaps in order
__global__ void __launch_bounds__(128) k_lu_swap(double* W, long long ld, long long c0, long long c1,
                                                 long long k0, int jb, const double* pivd) {
    __shared__ long long s_piv[kNb];
    if (threadIdx.x < kNb) s_piv[threadIdx.x] = threadIdx.x < jb ? (long long)pivd[threadIdx.x] : 0;
    __syncthreads();
    const long long c = c0 + (long long)blockIdx.x * 128 + threadIdx.x;
    if (c >= c1) return;
    double* col = W + c * ld;
    for (int j = 0; j < jb; ++j) {
        const long long p = s_piv[j];
        if (p != k0 + j) {
            const double t = col[k0 + j];
            col[k0 + j] = col[p];
            col[p] = t;
        }
    }
}

// ---- the FP64 tile kernel ---------------------------------------------------------------
// C (M x N) = A (M x 128) B (128 x N)      MODE 0   (C may be B itself when M <= 128: every
// C (M x N) -= A (M x 128) B (128 x N)     MODE 1    tile reads all of its B before it writes)
// all column-major; CTA tile 128 x 64, 128 threads, 8 x 8 outputs per thread: rows
// {2 tx, 2 tx + 1} + 32 i (i < 4), columns 8 ty .. 8 ty + 7; K in slabs of 16 through shared
// memory, the next slab's global loads in flight while the current one is multiplied.
// 64 DFMA per 8 shared-memory loads of 16 bytes: bound by the FP64 pipe.
constexpr int kGM = 128, kGN = 64, kGK = 16;

template <int MODE>
__global__ void __launch_bounds__(128, 2) k_lu_gemm(const double* __restrict__ Ap, long long lda,
                                                    const double* Bp, long long ldb, double* Cp,
                                                    long long ldc, long long M, long long N) {
    __shared__ __align__(16) double As[2][kGK][kGM];
    __shared__ __align__(16) double Bs[2][kGK][kGN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * kGM, n0 = (long long)blockIdx.y * kGN;
    const bool fullM = m0 + kGM <= M;

    // staging registers of one slab: A 128 x 16 -> 8 pairs per thread, B 16 x 64 -> 4 pairs
    double2 ra[8], rb[4];
    auto load_slab = [&](int kc) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = tid + 128 * q, kk = idx >> 6, m = (idx & 63) * 2;
            const double* src = Ap + (long long)(kc + kk) * lda + m0 + m;
            if (fullM || m0 + m + 1 < M)
                ra[q] = *reinterpret_cast<const double2*>(src);
            else
                ra[q] = make_double2(m0 + m < M ? src[0] : 0.0, 0.0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + 128 * q, col = idx >> 3, kk = (idx & 7) * 2;
            if (n0 + col < N)
                rb[q] = *reinterpret_cast<const double2*>(Bp + (n0 + col) * ldb + kc + kk);
            else
                rb[q] = make_double2(0.0, 0.0);
        }
    };
    auto store_slab = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = tid + 128 * q, kk = idx >> 6, m = (idx & 63) * 2;
            *reinterpret_cast<double2*>(&As[buf][kk][m]) = ra[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + 128 * q, col = idx >> 3, kk = (idx & 7) * 2;
            Bs[buf][kk][col] = rb[q].x;
            Bs[buf][kk + 1][col] = rb[q].y;
        }
    };

    // MODE 1: the accumulators start from C (64 independent loads per thread, in flight while
    // the first slab arrives) and the products are subtracted: C - sum a b, stored as is
    double acc[8][8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const long long col = n0 + 8 * ty + jj;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long m = m0 + 2 * tx + 32 * i;
            double2 v = make_double2(0.0, 0.0);
            if (MODE == 1 && col < N) {
                const double* c = Cp + col * ldc + m;
                if (m + 1 < M)
                    v = *reinterpret_cast<const double2*>(c);
                else if (m < M)
                    v.x = c[0];
            }
            acc[2 * i][jj] = v.x;
            acc[2 * i + 1][jj] = v.y;
        }
    }

    load_slab(0);
    store_slab(0);
    __syncthreads();
    for (int s = 0; s < kNb / kGK; ++s) {
        const int cur = s & 1;
        if (s + 1 < kNb / kGK) load_slab((s + 1) * kGK);
#pragma unroll
        for (int kk = 0; kk < kGK; ++kk) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(&As[cur][kk][2 * tx + 32 * i]);
                a[2 * i] = v.x;
                a[2 * i + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(&Bs[cur][kk][8 * ty + 2 * i]);
                b[2 * i] = v.x;
                b[2 * i + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    acc[i][jj] = (MODE == 1) ? fma(-a[i], b[jj], acc[i][jj]) : fma(a[i], b[jj], acc[i][jj]);
        }
        if (s + 1 < kNb / kGK) {
            store_slab(cur ^ 1);
            __syncthreads();
        }
    }
    // (MODE 0 with C == B: every thread of the CTA has finished reading B -- the last slab was
    // stored to shared memory before the barrier above)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const long long col = n0 + 8 * ty + jj;
        if (col >= N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long m = m0 + 2 * tx + 32 * i;
            double* c = Cp + col * ldc + m;
            if (m + 1 < M) {
                *reinterpret_cast<double2*>(c) = make_double2(acc[2 * i][jj], acc[2 * i + 1][jj]);
            } else if (m < M) {
                c[0] = acc[2 * i][jj];
            }
        }
    }
}

// ---- solve sweeps -----------------------------------------------------------------------
// v[k0 .. k0+jb) <- inv(T) v[k0 .. k0+jb), T = the triangle of the panel's diagonal block:
// LOWER unit (forward sweep) or upper with its diagonal (backward sweep).  One CTA.
template <bool LOWER>
__global__ void __launch_bounds__(kNb) k_lu_trsv(const double* W, long long ld, long long lc0,
                                                 long long k0, int jb, double* v) {
    __shared__ double s_v[kNb];
    const int i = threadIdx.x;
    if (i < jb) s_v[i] = v[k0 + i];
    __syncthreads();
    if (LOWER) {
        for (int c = 0; c < jb; ++c) {
            const double vc = s_v[c];
            __syncthreads();
            if (i > c && i < jb) s_v[i] = fma(-W[(lc0 + c) * ld + k0 + i], vc, s_v[i]);
            __syncthreads();
        }
    } else {
        for (int c = jb - 1; c >= 0; --c) {
            if (i == c) s_v[c] = s_v[c] / W[(lc0 + c) * ld + k0 + c];
            __syncthreads();
            const double vc = s_v[c];
            if (i < c) s_v[i] = fma(-W[(lc0 + c) * ld + k0 + i], vc, s_v[i]);
            __syncthreads();
        }
    }
    if (i < jb) v[k0 + i] = s_v[i];
}

// v[r] -= sum_c W(r, lc0 + c) y[c] for rows [r0, r1), y = v[k0 .. k0+jb): one thread per row,
// coalesced over the rows for every column
__global__ void __launch_bounds__(256) k_lu_gemv_sub(const double* W, long long ld, long long lc0,
                                                     long long k0, int jb, long long r0,
                                                     long long r1, double* v) {
    __shared__ double s_y[kNb];
    if (threadIdx.x < kNb) s_y[threadIdx.x] = threadIdx.x < jb ? v[k0 + threadIdx.x] : 0.0;
    __syncthreads();
    const long long r = r0 + (long long)blockIdx.x * 256 + threadIdx.x;
    if (r >= r1) return;
    const double* a = W + lc0 * ld + r;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int c = 0;
    for (; c + 3 < jb; c += 4) {
        acc0 = fma(a[(long long)c * ld], s_y[c], acc0);
        acc1 = fma(a[(long long)(c + 1) * ld], s_y[c + 1], acc1);
        acc2 = fma(a[(long long)(c + 2) * ld], s_y[c + 2], acc2);
        acc3 = fma(a[(long long)(c + 3) * ld], s_y[c + 3], acc3);
    }
    for (; c < jb; ++c) acc0 = fma(a[(long long)c * ld], s_y[c], acc0);
    v[r] -= (acc0 + acc1) + (acc2 + acc3);
}

// swaps of one panel on a vector (forward sweep), pivots from the persistent table
__global__ void k_lu_swap_vec(double* v, const long long* piv, long long k0, int jb) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int j = 0; j < jb; ++j) {
        const long long p = piv[k0 + j];
        if (p != k0 + j) {
            const double t = v[k0 + j];
            v[k0 + j] = v[p];
            v[p] = t;
        }
    }
}

__global__ void k_lu_store_piv(const double* pivd, long long* piv, long long k0, int jb) {
    const int j = threadIdx.x;
    if (j < jb) piv[k0 + j] = (long long)pivd[j];
}

// W(:, lc) *= scale[global column of lc]: rows of G -> columns of S = diag(s) G
__global__ void __launch_bounds__(256) k_lu_scale_cols(double* W, long long ld, long long n,
                                                       long long ncols, int P, int rank,
                                                       const double* scale) {
    const long long lc = blockIdx.x;   // (grid.x: up to 2^31 - 1 columns)
    if (lc >= ncols) return;
    const long long gc = ((lc / kNb) * P + rank) * kNb + (lc % kNb);
    const double s = scale[gc];
    double* col = W + lc * ld;
    for (long long r = threadIdx.x; r < n; r += 256) col[r] *= s;
}

__global__ void k_lu_mul(const double* a, const double* b, double* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = b ? a[i] * b[i] : a[i];
}

long long local_columns(long long n, int P, int rank) {
    const long long nblk = (n + kNb - 1) / kNb;
    long long cols = 0;
    for (long long kb = rank; kb < nblk; kb += P) cols += std::min<long long>(kNb, n - kb * kNb);
    return cols;
}

}  // namespace

void lu_free(icqb_ctx* ctx) {
    cudaFree(ctx->d_lu_piv);
    cudaFree(ctx->d_lu_buf);
    cudaFree(ctx->d_lu_small);
    ctx->d_lu_piv = nullptr;
    ctx->d_lu_buf = nullptr;
    ctx->d_lu_small = nullptr;
    ctx->lu_n = 0;
}

}  // namespace icqb

using namespace icqb;

extern "C" {

int64_t icqb_lu_local_columns(const icqb_ctx* ctx, int64_t n) {
    if (!ctx || n < 0) return -1;
    return local_columns(n, comm_nranks(ctx), comm_rank(ctx));
}

int icqb_lu_factor_dev(icqb_ctx* ctx, uint64_t dW, int64_t n, int64_t ld, uint64_t dscale,
                       int* info) {
    if (!ctx) return ICQB_EARG;
    if (!dW || n <= 0 || ld < n || (ld & 1))
        return set_error(ctx, ICQB_EARG, "icqb_lu_factor_dev: null matrix, ld < n or odd ld");
    if (dW & 15) return set_error(ctx, ICQB_EARG, "icqb_lu_factor_dev: the matrix must be 16-byte aligned");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double* W = reinterpret_cast<double*>(dW);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long nblk = (n + kNb - 1) / kNb;
    const long long ncols = local_columns(n, P, rank);

    // persistent: pivots; per factorisation: panel buffer, small state
    lu_free(ctx);
    const long long lpMax = (n + 1) & ~1LL;
    const size_t bufDoubles = (size_t)(kPackL21 + (long long)kNb * lpMax);
    ICQB_CUDA(ctx, cudaMalloc(&ctx->d_lu_piv, sizeof(long long) * (size_t)(n + kNb)));
    ICQB_CUDA(ctx, cudaMalloc(&ctx->d_lu_buf, sizeof(double) * bufDoubles));
    // small: rowbuf [4*128] | part_val [kMaxPart] | part_idx [kMaxPart] | piv [128] | pivval [128] | ticket, info
    const size_t smallDoubles = 4 * kNb + 2 * kMaxPart + 2 * kNb + 8;
    ICQB_CUDA(ctx, cudaMalloc(&ctx->d_lu_small, sizeof(double) * smallDoubles));
    ICQB_CUDA(ctx, cudaMemsetAsync(ctx->d_lu_small, 0, sizeof(double) * smallDoubles, st));
    ctx->lu_n = n;
    double* sm = ctx->d_lu_small;
    PanelArgs A;
    A.W = W;
    A.ld = ld;
    A.n = n;
    A.rowbuf = sm;
    A.part_val = sm + 4 * kNb;
    A.part_idx = reinterpret_cast<long long*>(sm + 4 * kNb + kMaxPart);
    A.piv = reinterpret_cast<long long*>(sm + 4 * kNb + 2 * kMaxPart);
    A.pivval = sm + 4 * kNb + 2 * kMaxPart + kNb;
    A.ticket = reinterpret_cast<unsigned int*>(sm + 4 * kNb + 2 * kMaxPart + 2 * kNb);
    A.info = reinterpret_cast<int*>(sm + 4 * kNb + 2 * kMaxPart + 2 * kNb + 2);
    double* buf = ctx->d_lu_buf;

    cudaEventRecord(ctx->ev0, st);
    if (dscale && ncols > 0) {
        k_lu_scale_cols<<<(unsigned)ncols, 256, 0, st>>>(W, ld, n, ncols, P, rank,
                                                                   reinterpret_cast<const double*>(dscale));
        ctx->launches += 1;
        ICQB_CUDA(ctx, cudaGetLastError());
    }
    ICQB_CUDA(ctx, cudaFuncSetAttribute(k_lu_trinv, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrinvSmem));

    for (long long kb = 0; kb < nblk; ++kb) {
        const int owner = (int)(kb % P);
        const long long k0 = kb * kNb;
        const int jb = (int)std::min<long long>(kNb, n - k0);
        const long long m2 = n - k0 - jb;                 // rows below the panel's diagonal block
        const long long lp = (m2 + 1) & ~1LL;
        const long long packed = kPackL21 + (long long)jb * lp;
        if (rank == owner) {
            A.lc0 = (kb / P) * kNb;
            A.k0 = k0;
            A.jb = jb;
            const long long rows = n - k0;
            const unsigned g0 = (unsigned)std::min<long long>((rows + 255) / 256, kMaxPart);
            k_lu_pivot0<<<g0, 256, 0, st>>>(A);
            ctx->launches += 1;
            for (int j = 0; j < jb; ++j) {
                const long long rem = n - (k0 + j + 1);
                if (rem <= 0) break;
                const long long g = (rem + 127) / 128;
                if (g > kMaxPart)
                    return set_error(ctx, ICQB_EARG, "icqb_lu_factor_dev: matrix too large for the panel kernels");
                k_lu_col<<<(unsigned)g, dim3(128, 4), 0, st>>>(A, j);
                ctx->launches += 1;
            }
            ICQB_CUDA(ctx, cudaGetLastError());
            if (m2 > 0) {
                k_lu_trinv<<<1, kNb, kTrinvSmem, st>>>(A, buf);
                k_lu_pack<<<(unsigned)std::min<long long>((m2 * jb + 255) / 256, 2048), 256, 0, st>>>(A, buf, lp);
                ctx->launches += 2;
            } else {
                k_lu_pack<<<1, 256, 0, st>>>(A, buf, lp);   // pivots only
                ctx->launches += 1;
            }
            ICQB_CUDA(ctx, cudaGetLastError());
        }
        if (P > 1) {
            int rc = comm_broadcast(ctx, buf, m2 > 0 ? packed : kNb, owner, st);
            if (rc != ICQB_OK) return rc;
        }
        k_lu_store_piv<<<1, kNb, 0, st>>>(buf, ctx->d_lu_piv, k0, jb);
        ctx->launches += 1;
        if (m2 <= 0) continue;
        // own columns right of the panel: local blocks lb with lb * P + rank > kb
        const long long lb0 = (kb >= rank) ? (kb - rank) / P + 1 : 0;
        const long long c0 = lb0 * kNb;
        if (c0 >= ncols) continue;
        const long long nc = ncols - c0;
        k_lu_swap<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(W, ld, c0, ncols, k0, jb, buf);
        // U12 = inv(L11) A12 (in place), then A22 -= L21 U12
        double* B = W + c0 * ld + k0;
        k_lu_gemm<0><<<dim3(1, (unsigned)((nc + kGN - 1) / kGN)), 128, 0, st>>>(
            buf + kPackInv, kNb, B, ld, B, ld, kNb, nc);
        k_lu_gemm<1><<<dim3((unsigned)((m2 + kGM - 1) / kGM), (unsigned)((nc + kGN - 1) / kGN)), 128, 0, st>>>(
            buf + kPackL21, lp, B, ld, B + kNb, ld, m2, nc);
        ctx->launches += 3;
        ICQB_CUDA(ctx, cudaGetLastError());
    }
    cudaEventRecord(ctx->ev1, st);
    int hinfo = 0;
    ICQB_CUDA(ctx, cudaMemcpyAsync(&hinfo, A.info, sizeof(int), cudaMemcpyDeviceToHost, st));
    ICQB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[kLu] = ms;
    if (P > 1) {
        // a zero pivot is seen by the panel's owner only: agree on the first one
        double* flag = ctx->d_lu_small;   // (rowbuf is free now)
        const double hv = hinfo > 0 ? -1.0 / hinfo : -2.0;   // max over ranks <-> smallest positive info
        ICQB_CUDA(ctx, cudaMemcpyAsync(flag, &hv, sizeof(double), cudaMemcpyHostToDevice, st));
        int rc = comm_allreduce_max(ctx, flag, 1, st);
        if (rc != ICQB_OK) return rc;
        double out = 0.0;
        ICQB_CUDA(ctx, cudaMemcpyAsync(&out, flag, sizeof(double), cudaMemcpyDeviceToHost, st));
        ICQB_CUDA(ctx, cudaStreamSynchronize(st));
        hinfo = out > -1.5 ? (int)llround(-1.0 / out) : 0;
    }
    if (info) *info = hinfo;
    return ICQB_OK;
}

int icqb_lu_solve_dev(icqb_ctx* ctx, uint64_t dW, int64_t n, int64_t ld, uint64_t dscale,
                      uint64_t db, uint64_t dx) {
    if (!ctx) return ICQB_EARG;
    if (!dW || !db || !dx || n <= 0 || ld < n)
        return set_error(ctx, ICQB_EARG, "icqb_lu_solve_dev: null pointer or bad sizes");
    if (!ctx->d_lu_piv || ctx->lu_n != n)
        return set_error(ctx, ICQB_ESTATE, "icqb_lu_solve_dev: no factorisation of this size (icqb_lu_factor_dev)");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const double* W = reinterpret_cast<const double*>(dW);
    double* v = reinterpret_cast<double*>(dx);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long nblk = (n + kNb - 1) / kNb;
    cudaEventRecord(ctx->ev0, st);
    // v = diag(s) b
    k_lu_mul<<<128, 256, 0, st>>>(reinterpret_cast<const double*>(db),
                                  reinterpret_cast<const double*>(dscale), v, n);
    ctx->launches += 1;
    // forward: replay the panels (swaps, L11, L21); the vector travels with the owner
    for (long long kb = 0; kb < nblk; ++kb) {
        const int owner = (int)(kb % P);
        const long long k0 = kb * kNb;
        const int jb = (int)std::min<long long>(kNb, n - k0);
        if (rank == owner) {
            const long long lc0 = (kb / P) * kNb;
            k_lu_swap_vec<<<1, 32, 0, st>>>(v, ctx->d_lu_piv, k0, jb);
            k_lu_trsv<true><<<1, kNb, 0, st>>>(W, ld, lc0, k0, jb, v);
            ctx->launches += 2;
            const long long r0 = k0 + jb;
            if (r0 < n) {
                k_lu_gemv_sub<<<(unsigned)((n - r0 + 255) / 256), 256, 0, st>>>(W, ld, lc0, k0, jb, r0, n, v);
                ctx->launches += 1;
            }
        }
        if (P > 1) {
            int rc = comm_broadcast(ctx, v + k0, n - k0, owner, st);
            if (rc != ICQB_OK) return rc;
        }
    }
    // backward: U
    for (long long kb = nblk - 1; kb >= 0; --kb) {
        const int owner = (int)(kb % P);
        const long long k0 = kb * kNb;
        const int jb = (int)std::min<long long>(kNb, n - k0);
        if (rank == owner) {
            const long long lc0 = (kb / P) * kNb;
            k_lu_trsv<false><<<1, kNb, 0, st>>>(W, ld, lc0, k0, jb, v);
            ctx->launches += 1;
            if (k0 > 0) {
                k_lu_gemv_sub<<<(unsigned)((k0 + 255) / 256), 256, 0, st>>>(W, ld, lc0, k0, jb, 0, k0, v);
                ctx->launches += 1;
            }
        }
        if (P > 1) {
            int rc = comm_broadcast(ctx, v, k0 + jb, owner, st);
            if (rc != ICQB_OK) return rc;
        }
    }
    cudaEventRecord(ctx->ev1, st);
    ICQB_CUDA(ctx, cudaGetLastError());
    ICQB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[kLuSolve] = ms;
    return ICQB_OK;
}

}  // extern "C"
