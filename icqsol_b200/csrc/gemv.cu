// y = A x (optionally row-scaled) for a row-major fp64 matrix: the HBM-bound
// kernel behind PoissonSolver.computeResponseField (numpy.dot,
// bem/icqPoissonSolver.py:36) and every CG iteration
// (solvers/icqConjugateGradient.py:57,66,74).
//
// Algorithmic traffic: 8*nrows*ncols bytes, streamed exactly once.
// Mapping: a CTA of 256 threads owns kRows consecutive rows; the threads stride
// the columns with 16-byte streaming loads (ld.global.nc.L1::no_allocate) of A,
// two column steps in flight per row (8 independent 16 B loads per thread), x is
// read through L1/L2 (it is shared by every CTA and stays resident); partial sums
// are combined with warp shuffles and one shared-memory pass -- a fixed order, so
// the result is reproducible run to run.  No tensor cores: one multiply-add per
// 8 bytes streamed.
#include "icqb_internal.cuh"

namespace icqb {

namespace {

constexpr int kThreads = 256;
constexpr int kRows = 4;

#ifdef ICQB_EMU   // CPU emulation build of the test suite (tools/emu): no PTX
__device__ __forceinline__ double2 ld_stream(const double* p) {
    return *reinterpret_cast<const double2*>(p);
}
#else
__device__ __forceinline__ double2 ld_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}
#endif

template <bool ALIGNED>
__global__ void __launch_bounds__(kThreads) k_gemv(const double* __restrict__ A, long long nrows,
                                                   long long ncols, long long ld,
                                                   const double* __restrict__ x,
                                                   double* __restrict__ y,
                                                   const double* __restrict__ rowscale,
                                                   const int* __restrict__ skip) {
    __shared__ double s_part[kRows][kThreads / 32];
    if (skip && *skip) return;
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * kRows;
    const double* a[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        long long row = row0 + r;
        if (row >= nrows) row = nrows - 1;  // clamp: duplicates work, never stored
        a[r] = A + row * ld;
    }
    double acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0;

    if (ALIGNED) {
        // columns in pairs; ncols2 pairs, plus an odd tail column
        const long long npair = ncols >> 1;
        long long c = tid;
        for (; c + kThreads < npair; c += 2 * kThreads) {
            const double2 x0 = __ldg(reinterpret_cast<const double2*>(x) + c);
            const double2 x1 = __ldg(reinterpret_cast<const double2*>(x) + c + kThreads);
            double2 v0[kRows], v1[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                v0[r] = ld_stream(a[r] + 2 * c);
                v1[r] = ld_stream(a[r] + 2 * (c + kThreads));
            }
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                acc[r] = fma(v0[r].x, x0.x, acc[r]);
                acc[r] = fma(v0[r].y, x0.y, acc[r]);
                acc[r] = fma(v1[r].x, x1.x, acc[r]);
                acc[r] = fma(v1[r].y, x1.y, acc[r]);
            }
        }
        for (; c < npair; c += kThreads) {
            const double2 x0 = __ldg(reinterpret_cast<const double2*>(x) + c);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const double2 v = ld_stream(a[r] + 2 * c);
                acc[r] = fma(v.x, x0.x, acc[r]);
                acc[r] = fma(v.y, x0.y, acc[r]);
            }
        }
        if ((ncols & 1) && tid == 0) {
            const double xl = x[ncols - 1];
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] = fma(a[r][ncols - 1], xl, acc[r]);
        }
    } else {
        for (long long c = tid; c < ncols; c += kThreads) {
            const double xv = __ldg(x + c);
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] = fma(__ldg(a[r] + c), xv, acc[r]);
        }
    }

#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        double v = acc[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) s_part[r][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < kRows) {
        const long long row = row0 + tid;
        if (row < nrows) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) v += s_part[tid][w];
            if (rowscale) v *= rowscale[row];
            y[row] = v;
        }
    }
}

}  // namespace

int launch_gemv(icqb_ctx* ctx, const double* dA, int64_t nrows, int64_t ncols, int64_t ld,
                const double* dx, double* dy, const double* drowscale, cudaStream_t stream,
                const int* skip) {
    if (nrows <= 0) return ICQB_OK;
    if (ncols < 0 || ld < ncols) return set_error(ctx, ICQB_EARG, "gemv: need ld >= ncols >= 0");
    const unsigned blocks = (unsigned)((nrows + kRows - 1) / kRows);
    const bool aligned = ((reinterpret_cast<uintptr_t>(dA) & 15) == 0) && ((ld & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);
    if (aligned)
        k_gemv<true><<<blocks, kThreads, 0, stream>>>(dA, nrows, ncols, ld, dx, dy, drowscale, skip);
    else
        k_gemv<false><<<blocks, kThreads, 0, stream>>>(dA, nrows, ncols, ld, dx, dy, drowscale, skip);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_gemv launch");
}

}  // namespace icqb
