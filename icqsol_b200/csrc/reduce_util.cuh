// Fixed-order reductions shared by the symmetric product and the conjugate gradient.
//
// Dot products of the CG (solvers/icqConjugateGradient.py:60-72: numpy.dot) are formed
// as one partial per CTA; the CTA that finishes last (atomic ticket) adds the partials in
// an order that depends only on their count, so a result is reproducible run to run and
// identical on every rank.
#pragma once

#include <cuda_runtime.h>

namespace icqb {

// Sum (or max) of 128 per-thread values held by the first four warps of a CTA
// (threads 0..127 in linear order).  s4 is a 4-entry shared array; the result is valid on
// thread 0 after the caller's __syncthreads().  Call from all threads of those warps.
template <bool MAX>
__device__ __forceinline__ void warp_part(double v, double* s4, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, off);
        v = MAX ? fmax(v, o) : v + o;
    }
    if ((tid & 31) == 0) s4[tid >> 5] = v;
}

template <bool MAX>
__device__ __forceinline__ double join4(const double* s4) {
    return MAX ? fmax(fmax(s4[0], s4[1]), fmax(s4[2], s4[3])) : (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

// Whole-CTA reduction of `count` partials from global memory (written by other CTAs
// before their ticket): thread t takes partials t, t+T, ... in order, then a binary tree
// over the T threads.  T = blockDim (power of two); s_buf holds T doubles.
// Every thread of the CTA must call; the result is returned on all threads.
template <bool MAX>
__device__ __forceinline__ double final_reduce(const double* part, int count, double* s_buf,
                                               int tid, int nthreads) {
    const volatile double* vp = part;
    double acc = 0.0;
    for (int q = tid; q < count; q += nthreads) {
        const double o = vp[q];
        acc = MAX ? fmax(acc, o) : acc + o;
    }
    s_buf[tid] = acc;
    __syncthreads();
    for (int off = nthreads >> 1; off > 0; off >>= 1) {
        if (tid < off) {
            const double o = s_buf[tid + off];
            s_buf[tid] = MAX ? fmax(s_buf[tid], o) : s_buf[tid] + o;
        }
        __syncthreads();
    }
    const double out = s_buf[0];
    __syncthreads();
    return out;
}

// Three reductions at once (sum, sum, max) -- the closing step of a CG iteration needs
// rho, |r|^2 and max|r| and should pay the tree's barriers once.  s_buf holds 3*T doubles.
__device__ __forceinline__ void final_reduce3(const double* sumA, const double* sumB,
                                              const double* maxC, int count, double* s_buf,
                                              int tid, int nthreads, double* outA, double* outB,
                                              double* outC) {
    const volatile double* pa = sumA;
    const volatile double* pb = sumB;
    const volatile double* pc = maxC;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int q = tid; q < count; q += nthreads) {
        a += pa[q];
        b += pb[q];
        c = fmax(c, pc[q]);
    }
    double* sa = s_buf;
    double* sb = s_buf + nthreads;
    double* sc = s_buf + 2 * nthreads;
    sa[tid] = a;
    sb[tid] = b;
    sc[tid] = c;
    __syncthreads();
    for (int off = nthreads >> 1; off > 0; off >>= 1) {
        if (tid < off) {
            sa[tid] += sa[tid + off];
            sb[tid] += sb[tid + off];
            sc[tid] = fmax(sc[tid], sc[tid + off]);
        }
        __syncthreads();
    }
    *outA = sa[0];
    *outB = sb[0];
    *outC = sc[0];
    __syncthreads();
}

// Ticket of the "last CTA" pattern: thread 0 publishes this CTA's partials (already
// stored) and learns whether every other CTA has published too.
__device__ __forceinline__ bool take_ticket(unsigned int* ticket, unsigned int nblocks) {
    __threadfence();
    return atomicAdd(ticket, 1u) == nblocks - 1;
}

}  // namespace icqb
