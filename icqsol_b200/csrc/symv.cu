// Matrix-vector product that streams only the stored half of the Green matrix.
//
// diag(A) G is symmetric (G[j,i] = G[i,j] A_i/A_j, the mirror rule of
// bem/icqLaplaceMatrices.cpp:97-98), so a product with G -- numpy.dot at
// bem/icqPoissonSolver.py:36 and the GEMV of every CG iteration,
// solvers/icqConjugateGradient.py:66 -- needs each pair entry only once.  In the
// block-lower-triangular storage (tiles J <= I of the locally owned row blocks,
// diagonal tiles complete) every off-diagonal tile G_IJ contributes
//      row part   r_I += G_IJ x_J            (as stored)
//      col part   c_J += G_IJ^T (gamma_I . x_I)   (the mirrored entries)
// and the result is  z = alpha . r + beta . c  with per-row vectors:
//      CG on diag(A) G:  gamma = A, alpha = A, beta = 1
//      plain G x:        gamma = A, alpha = 1, beta = 1/A
// Algorithmic traffic: 8 bytes per STORED entry = 4 N^2 + O(128 N) per product,
// half of the row-major GEMV.
//
// Mapping (k_symv): the stored tiles of a rank are ONE contiguous array, cut into
// 8 KB chunks (8 rows x 128 columns of a tile).  A persistent grid of one CTA per SM,
// 8 warps each, divides the chunk sequence evenly: every warp owns one contiguous
// span, so there is no tail and no tile-size granularity.  Each warp is its own
// pipeline: lane 0 keeps kDepth bulk copies (cp.async.bulk ... mbarrier::complete_tx,
// SASS UBLKCP, L2 evict-first) in flight into the warp's private shared-memory ring
// -- 192 KB of loads outstanding per SM without holding a register -- and the 32
// lanes consume a landed chunk from shared memory: lane l owns columns
// {2l, 2l+1, 64+2l, 65+2l} (conflict-free 16-byte reads), column sums stay in its
// registers for the whole tile, the eight row sums of a chunk are reduced across the
// lanes with a transposed butterfly (9 shuffles per 8 rows instead of 40).
// Per tile a warp stores 128 row partials; column partials go to the tile's slot when
// the warp began the tile and to a per-warp "head" slot when its span starts inside the
// tile (the host knows the partition and tells the reduce kernel which head slots belong
// to which tile).  k_symv_reduce adds the partials per global index in a fixed order
// (no atomics: results are reproducible) and applies alpha/beta.  Inside the CG it also
// forms the new search direction and p.(A p) (SymvFuse, icqb_internal.cuh).
#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"
#include "reduce_util.cuh"

namespace icqb {

namespace {

constexpr int kT = 128;                               // tile edge
constexpr int kChunkRows = 8;
constexpr int kChunksPerTile = kT / kChunkRows;       // 16
constexpr int kChunkElems = kChunkRows * kT;          // 1024 doubles
constexpr int kChunkBytes = kChunkElems * 8;          // 8 KB
constexpr int kWarps = 8;                             // independent pipelines per CTA
constexpr int kDepth = 3;                             // chunks in flight per warp
constexpr int kThreads = kWarps * 32;
constexpr int kRedGroups = 8;                         // thread groups of the reduce kernel
constexpr int kRedThreads = kT * kRedGroups;          // 1024
static_assert(kRedGroups == 8, "k_symv_reduce joins exactly eight groups");
constexpr int kSmemBytes =
    kWarps * kDepth * kChunkBytes + kWarps * kT * 8 + kWarps * kDepth * 8;   // 204,992 B

// Optional float32 product (k_symv_f32, for the mixed-precision CG; off by default): the same chunks of 8 rows x 128 columns are 4 KB,
// the ring holds twice as many of them
constexpr int kDepth32 = 6;
constexpr int kChunkBytes32 = kChunkElems * 4;
constexpr int kSmemBytes32 =
    kWarps * kDepth32 * kChunkBytes32 + kWarps * kT * 8 + kWarps * kDepth32 * 8;   // 205,184 B

struct SymvTask {
    int lt;        // local strip (tile row) index
    int J;         // global source tile, J <= I
};

struct SymvParams {
    const double* A;            // tile-contiguous lower storage (LowerLayout)
    long long n;
    LowerLayout L;
    int nTasks;                 // == number of stored tiles; task t describes tile t
    const SymvTask* tasks;
    const double* x;            // [n]
    const double* gamma;        // [n] or null
    const double* alpha;        // [n] or null
    const double* beta;         // [n] or null
    double* Wc;                 // [nLocal][wcols] column partials of the warp that began a tile
    double* Wr;                 // [nTasks][128]  row partials per tile
    double* Wh;                 // [head slots][128] column partials of spans starting inside a tile
    const int* headSlot;        // [warps] slot of the warp's head segment, -1: none
    const int* headColBegin;    // [column tiles + 1] CSR over column tiles J ...
    const int* headColSlots;    // ... of the head slots whose tile lies in column tile J
    long long wcols;
    double* z;                  // [n]
    const int* skip;            // when non-null and *skip != 0 the kernels return at once
    SymvFuse F;                 // CG fusion (all null: plain product)
    SymvPeer X;                 // peer exchange (k_symv_reduce<true> only)
    const float* A32 = nullptr; // k_symv_f32 (optional): float32 copy of A, same tile layout
};

#ifdef ICQB_EMU   // CPU emulation build of the test suite (tools/emu): no PTX
// mbarrier modelled as the count of completed phases; a bulk copy completes when issued
__device__ __forceinline__ void mbar_init(uint64_t* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((*reinterpret_cast<volatile uint64_t*>(bar) & 1) == parity) emu::yield();
}
__device__ __forceinline__ uint64_t policy_evict_first() { return 0; }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t) {
    memcpy(dst, src, bytes);
    *bar += 1;
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared bulk copy, completion counted in bytes on `bar`; the matrix is read
// once per product, so its lines are marked evict-first in L2 (the vectors, partial
// arrays and the preconditioner blocks of the CG stay resident)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

#endif

// entry c of the multiplied vector: x, or the CG direction w + beta * p_old
__device__ __forceinline__ double vec_at(const SymvParams& P, double cgBeta, long long c) {
    const double xv = __ldg(P.x + c);
    return P.F.w ? fma(cgBeta, xv, __ldg(P.F.w + c)) : xv;
}

__global__ void __launch_bounds__(kThreads, 1) k_symv(const SymvParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (P.skip && *P.skip) return;
    double* s_ring = reinterpret_cast<double*>(smem_raw);            // [kWarps][kDepth][kChunkElems]
    double* s_mul = s_ring + kWarps * kDepth * kChunkElems;          // [kWarps][kT]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_mul + kWarps * kT);   // [kWarps][kDepth]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * kWarps + warp;
    const long long nW = (long long)gridDim.x * kWarps;
    const long long total = (long long)P.nTasks * kChunksPerTile;
    const long long c0 = total * gw / nW, c1 = total * (gw + 1) / nW;   // this warp's chunks
    const long long nCh = c1 - c0;
    if (nCh <= 0) return;

    double* ring = s_ring + warp * kDepth * kChunkElems;
    double* mul = s_mul + warp * kT;
    uint64_t* bar = s_bar + warp * kDepth;
    uint64_t policy = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kDepth; ++s) mbar_init(&bar[s], 1);
#ifndef ICQB_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        policy = policy_evict_first();
    }
    __syncwarp();
    if (lane == 0) {
        const int pre = (int)min((long long)kDepth, nCh);
        for (int s = 0; s < pre; ++s) {
            mbar_expect_tx(&bar[s], kChunkBytes);
            bulk_g2s(ring + s * kChunkElems, P.A + (c0 + s) * kChunkElems, kChunkBytes, &bar[s],
                     policy);
        }
    }

    const double cgBeta = P.F.w ? *P.F.beta : 0.0;

    // state of the tile being processed
    long long curTile = -1;
    bool diagTile = false, ragged = false, ownsStart = false;
    bool v0 = true, v1 = true, v2 = true, v3 = true;
    int nrows = 0, curLt = 0;
    long long colA = 0;
    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    double ca0 = 0, ca1 = 0, ca2 = 0, ca3 = 0;

    // column partials of the finished segment: to the tile's slot when this warp began
    // the tile, else to the warp's head slot
    auto flush = [&]() {
        if (curTile < 0 || diagTile) return;
        double* wc = ownsStart ? P.Wc + (long long)curLt * P.wcols + colA
                               : P.Wh + (long long)__ldg(P.headSlot + gw) * kT + 2 * lane;
        reinterpret_cast<double2*>(wc)[0] = make_double2(ca0, ca1);
        reinterpret_cast<double2*>(wc + 64)[0] = make_double2(ca2, ca3);
    };

    for (long long i = 0; i < nCh; ++i) {
        const long long ch = c0 + i;
        const long long tile = ch / kChunksPerTile;
        const int rbase = (int)(ch % kChunksPerTile) * kChunkRows;

        if (tile != curTile) {
            flush();
            curTile = tile;
            ownsStart = (tile * kChunksPerTile >= c0);
            const SymvTask task = P.tasks[tile];
            curLt = task.lt;
            const long long obs0 = P.L.obs0(task.lt), obsEnd = P.L.obsEnd(task.lt);
            const long long I = obs0 / kT, J = task.J;
            nrows = (int)min((long long)kT, obsEnd - obs0);
            diagTile = (J == I);
            // valid columns: tiles below the diagonal tile are complete; the diagonal tile ends at n
            const long long cEnd = diagTile ? min(P.n, (J + 1) * kT) : (J + 1) * kT;
            ragged = cEnd < (J + 1) * kT;
            colA = J * kT + 2 * lane;
            const long long colB = colA + 64;
            v0 = colA < cEnd;
            v1 = colA + 1 < cEnd;
            v2 = colB < cEnd;
            v3 = colB + 1 < cEnd;
            x0 = v0 ? vec_at(P, cgBeta, colA) : 0.0;
            x1 = v1 ? vec_at(P, cgBeta, colA + 1) : 0.0;
            x2 = v2 ? vec_at(P, cgBeta, colB) : 0.0;
            x3 = v3 ? vec_at(P, cgBeta, colB + 1) : 0.0;
            ca0 = ca1 = ca2 = ca3 = 0.0;
            if (!diagTile) {
                // multipliers of the mirrored entries: gamma_i x_i for the tile's 128 rows
                __syncwarp();
#pragma unroll
                for (int q = 0; q < kT / 32; ++q) {
                    const int r = q * 32 + lane;
                    double m = 0.0;
                    if (r < nrows) {
                        m = vec_at(P, cgBeta, obs0 + r);
                        if (P.gamma) m *= __ldg(P.gamma + obs0 + r);
                    }
                    mul[r] = m;
                }
                __syncwarp();
            }
        }

        const int s = (int)(i % kDepth);
        mbar_wait(&bar[s], (uint32_t)((i / kDepth) & 1));
        const double* sp = ring + s * kChunkElems + 2 * lane;

        double part[kChunkRows];
#pragma unroll
        for (int u = 0; u < kChunkRows; ++u) {
            const int r = rbase + u;
            part[u] = 0.0;
            if (r < nrows) {   // uniform across the warp
                double2 a01 = *reinterpret_cast<const double2*>(sp + u * kT);
                double2 a23 = *reinterpret_cast<const double2*>(sp + u * kT + 64);
                if (ragged) {   // columns beyond n were never written: mask them
                    if (!v0) a01.x = 0.0;
                    if (!v1) a01.y = 0.0;
                    if (!v2) a23.x = 0.0;
                    if (!v3) a23.y = 0.0;
                }
                part[u] = fma(a23.y, x3, fma(a23.x, x2, fma(a01.y, x1, a01.x * x0)));
                if (!diagTile) {
                    const double m = mul[r];   // same address on every lane: broadcast
                    ca0 = fma(a01.x, m, ca0);
                    ca1 = fma(a01.y, m, ca1);
                    ca2 = fma(a23.x, m, ca2);
                    ca3 = fma(a23.y, m, ca3);
                }
            }
        }
        // every lane has read its part of the stage: lane 0 may refill it
        __syncwarp();
        if (lane == 0 && i + kDepth < nCh) {
            mbar_expect_tx(&bar[s], kChunkBytes);
            bulk_g2s(ring + s * kChunkElems, P.A + (ch + kDepth) * kChunkElems, kChunkBytes, &bar[s],
                     policy);
        }

        // transposed butterfly: 8 row sums over 32 lanes in 4+2+1+1+1 shuffles; row
        // (lane >> 2) of the chunk ends up on the four lanes of its group
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double send = b4 ? part[q] : part[q + 4];
            const double keep = b4 ? part[q + 4] : part[q];
            part[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double send = b3 ? part[q] : part[q + 2];
            const double keep = b3 ? part[q + 2] : part[q];
            part[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const double send = b2 ? part[0] : part[1];
            const double keep = b2 ? part[1] : part[0];
            part[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        part[0] += __shfl_xor_sync(0xffffffffu, part[0], 2);
        part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
        if ((lane & 3) == 0) P.Wr[tile * kT + rbase + (lane >> 2)] = part[0];
    }
    flush();
}

// Optional (B200: 4.6 TB/s of float32 entries, tests/test_gpu_parity.py::test_float32_product_kernel): the same product on a
// FLOAT32 copy of the stored tiles -- half the bytes -- for the late iterations of the CG, where
// the direction is so small that a product accurate to 6e-8 of |A||p| still keeps the residual
// recurrence honest to well below the stopping threshold (inexact Krylov;
// tools/experiments/pcg_mixed_precision.py: -25..28 % traffic at an unchanged 5e-13 backward
// error).  Same static partition, chunks (8 rows x 128 columns = 4 KB), row butterfly, partial
// arrays and reduce kernel as k_symv; a lane owns four consecutive columns (one 16-byte
// read), entries are widened to double before they meet the double vectors, all sums in double.
__global__ void __launch_bounds__(kThreads, 1) k_symv_f32(const SymvParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (P.skip && *P.skip) return;
    float* s_ring = reinterpret_cast<float*>(smem_raw);              // [kWarps][kDepth32][kChunkElems]
    double* s_mul = reinterpret_cast<double*>(s_ring + kWarps * kDepth32 * kChunkElems);   // [kWarps][kT]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_mul + kWarps * kT);   // [kWarps][kDepth32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * kWarps + warp;
    const long long nW = (long long)gridDim.x * kWarps;
    const long long total = (long long)P.nTasks * kChunksPerTile;
    const long long c0 = total * gw / nW, c1 = total * (gw + 1) / nW;   // this warp's chunks
    const long long nCh = c1 - c0;
    if (nCh <= 0) return;

    float* ring = s_ring + warp * kDepth32 * kChunkElems;
    double* mul = s_mul + warp * kT;
    uint64_t* bar = s_bar + warp * kDepth32;
    uint64_t policy = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kDepth32; ++s) mbar_init(&bar[s], 1);
#ifndef ICQB_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        policy = policy_evict_first();
    }
    __syncwarp();
    if (lane == 0) {
        const int pre = (int)min((long long)kDepth32, nCh);
        for (int s = 0; s < pre; ++s) {
            mbar_expect_tx(&bar[s], kChunkBytes32);
            bulk_g2s(ring + s * kChunkElems, P.A32 + (c0 + s) * kChunkElems, kChunkBytes32, &bar[s],
                     policy);
        }
    }

    const double cgBeta = P.F.w ? *P.F.beta : 0.0;

    // state of the tile being processed
    long long curTile = -1;
    bool diagTile = false, ragged = false, ownsStart = false;
    bool v0 = true, v1 = true, v2 = true, v3 = true;
    int nrows = 0, curLt = 0;
    long long colA = 0;
    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    double ca0 = 0, ca1 = 0, ca2 = 0, ca3 = 0;

    // column partials of the finished segment: to the tile's slot when this warp began
    // the tile, else to the warp's head slot
    auto flush = [&]() {
        if (curTile < 0 || diagTile) return;
        double* wc = ownsStart ? P.Wc + (long long)curLt * P.wcols + colA
                               : P.Wh + (long long)__ldg(P.headSlot + gw) * kT + 4 * lane;
        reinterpret_cast<double2*>(wc)[0] = make_double2(ca0, ca1);
        reinterpret_cast<double2*>(wc + 2)[0] = make_double2(ca2, ca3);
    };

    for (long long i = 0; i < nCh; ++i) {
        const long long ch = c0 + i;
        const long long tile = ch / kChunksPerTile;
        const int rbase = (int)(ch % kChunksPerTile) * kChunkRows;

        if (tile != curTile) {
            flush();
            curTile = tile;
            ownsStart = (tile * kChunksPerTile >= c0);
            const SymvTask task = P.tasks[tile];
            curLt = task.lt;
            const long long obs0 = P.L.obs0(task.lt), obsEnd = P.L.obsEnd(task.lt);
            const long long I = obs0 / kT, J = task.J;
            nrows = (int)min((long long)kT, obsEnd - obs0);
            diagTile = (J == I);
            // valid columns: tiles below the diagonal tile are complete; the diagonal tile ends at n
            const long long cEnd = diagTile ? min(P.n, (J + 1) * kT) : (J + 1) * kT;
            ragged = cEnd < (J + 1) * kT;
            colA = J * kT + 4 * lane;     // this lane's four consecutive columns
            v0 = colA < cEnd;
            v1 = colA + 1 < cEnd;
            v2 = colA + 2 < cEnd;
            v3 = colA + 3 < cEnd;
            x0 = v0 ? vec_at(P, cgBeta, colA) : 0.0;
            x1 = v1 ? vec_at(P, cgBeta, colA + 1) : 0.0;
            x2 = v2 ? vec_at(P, cgBeta, colA + 2) : 0.0;
            x3 = v3 ? vec_at(P, cgBeta, colA + 3) : 0.0;
            ca0 = ca1 = ca2 = ca3 = 0.0;
            if (!diagTile) {
                // multipliers of the mirrored entries: gamma_i x_i for the tile's 128 rows
                __syncwarp();
#pragma unroll
                for (int q = 0; q < kT / 32; ++q) {
                    const int r = q * 32 + lane;
                    double m = 0.0;
                    if (r < nrows) {
                        m = vec_at(P, cgBeta, obs0 + r);
                        if (P.gamma) m *= __ldg(P.gamma + obs0 + r);
                    }
                    mul[r] = m;
                }
                __syncwarp();
            }
        }

        const int s = (int)(i % kDepth32);
        mbar_wait(&bar[s], (uint32_t)((i / kDepth32) & 1));
        const float* sp = ring + s * kChunkElems + 4 * lane;

        double part[kChunkRows];
        if (!ragged && !diagTile && rbase + kChunkRows <= nrows) {
            // the common case -- a complete chunk of a tile below the diagonal: no masks, no
            // per-row tests (half the instructions of the general loop; this kernel moves half
            // the bytes per chunk, so the instruction count per byte matters twice as much)
#pragma unroll
            for (int u = 0; u < kChunkRows; ++u) {
                const float4 af = *reinterpret_cast<const float4*>(sp + u * kT);   // 16 bytes
                const double a0 = (double)af.x, a1 = (double)af.y, a2 = (double)af.z, a3 = (double)af.w;
                part[u] = fma(a3, x3, fma(a2, x2, fma(a1, x1, a0 * x0)));
                const double m = mul[rbase + u];   // same address on every lane: broadcast
                ca0 = fma(a0, m, ca0);
                ca1 = fma(a1, m, ca1);
                ca2 = fma(a2, m, ca2);
                ca3 = fma(a3, m, ca3);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kChunkRows; ++u) {
                const int r = rbase + u;
                part[u] = 0.0;
                if (r < nrows) {   // uniform across the warp
                    const float4 af = *reinterpret_cast<const float4*>(sp + u * kT);   // 16 bytes
                    double2 a01 = make_double2((double)af.x, (double)af.y);
                    double2 a23 = make_double2((double)af.z, (double)af.w);
                    if (ragged) {   // columns beyond n were never written: mask them
                        if (!v0) a01.x = 0.0;
                        if (!v1) a01.y = 0.0;
                        if (!v2) a23.x = 0.0;
                        if (!v3) a23.y = 0.0;
                    }
                    part[u] = fma(a23.y, x3, fma(a23.x, x2, fma(a01.y, x1, a01.x * x0)));
                    if (!diagTile) {
                        const double m = mul[r];   // same address on every lane: broadcast
                        ca0 = fma(a01.x, m, ca0);
                        ca1 = fma(a01.y, m, ca1);
                        ca2 = fma(a23.x, m, ca2);
                        ca3 = fma(a23.y, m, ca3);
                    }
                }
            }
        }
        // every lane has read its part of the stage: lane 0 may refill it
        __syncwarp();
        if (lane == 0 && i + kDepth32 < nCh) {
            mbar_expect_tx(&bar[s], kChunkBytes32);
            bulk_g2s(ring + s * kChunkElems, P.A32 + (ch + kDepth32) * kChunkElems, kChunkBytes32, &bar[s],
                     policy);
        }

        // transposed butterfly: 8 row sums over 32 lanes in 4+2+1+1+1 shuffles; row
        // (lane >> 2) of the chunk ends up on the four lanes of its group
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double send = b4 ? part[q] : part[q + 4];
            const double keep = b4 ? part[q + 4] : part[q];
            part[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double send = b3 ? part[q] : part[q + 2];
            const double keep = b3 ? part[q + 2] : part[q];
            part[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const double send = b2 ? part[0] : part[1];
            const double keep = b2 ? part[1] : part[0];
            part[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        part[0] += __shfl_xor_sync(0xffffffffu, part[0], 2);
        part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
        if ((lane & 3) == 0) P.Wr[tile * kT + rbase + (lane >> 2)] = part[0];
    }
    flush();
}

// z[j] = alpha_j * (row partials of j's strip over its tiles) + beta_j * (column partials of
// the strips below).  One CTA per 128 global indices, kRedGroups thread groups split the two
// sums (many independent loads in flight: the kernel is latency bound) and are combined in a
// fixed order.  With SymvFuse the CTA also stores the new direction
// p_j = w_j + beta p_j and leaves its share of p.z; the last CTA adds the shares.
// PEER (optional peer exchange, validated on 4 and 8 B200): the partial product and its p.z also go to slot (parity, rank) of
// every rank's exchange buffer, and the last CTA raises this rank's flag on every rank.
template <bool PEER>
__global__ void __launch_bounds__(kRedThreads) k_symv_reduce(const SymvParams P) {
    __shared__ double s_row[kRedGroups][kT], s_col[kRedGroups][kT];
    __shared__ double s_dot[4];
    __shared__ double s_fin[kRedThreads];
    __shared__ int s_last;
    if (P.skip && *P.skip) return;
    const int T = blockIdx.x, c = threadIdx.x, g = threadIdx.y;
    const int tid = g * kT + c;
    const long long j = (long long)T * kT + c;
    const long long jl = min(j, P.n - 1);       // clamp: lanes beyond n never store
    double col = 0.0, row = 0.0;
    // strips strictly below tile row T, first block then second block
    for (int blk = 0; blk < 2; ++blk) {
        const long long b0 = blk == 0 ? P.L.r0 : P.L.q0;
        const long long ltBeg = blk == 0 ? 0 : P.L.nTiles0;
        const long long ltEnd = blk == 0 ? P.L.nTiles0 : P.L.nLocal();
        const long long first = max(0LL, (long long)T + 1 - b0 / kT);
#pragma unroll 4
        for (long long lt = ltBeg + first + g; lt < ltEnd; lt += kRedGroups)
            col += P.Wc[lt * P.wcols + jl];
    }
    // column partials of the spans that started inside a tile of this column tile
    if (P.headColBegin) {
        const int hEnd = __ldg(P.headColBegin + T + 1);
        for (int q = __ldg(P.headColBegin + T) + g; q < hEnd; q += kRedGroups)
            col += P.Wh[(long long)__ldg(P.headColSlots + q) * kT + c];
    }
    // this tile row, if it is stored here
    const int lt = P.L.stripOf((long long)T * kT);
    if (lt >= 0) {
        const long long base = P.L.tileBase(lt);
#pragma unroll 4
        for (int Jt = g; Jt <= T; Jt += kRedGroups) row += P.Wr[(base + Jt) * kT + c];
    }
    s_row[g][c] = row;
    s_col[g][c] = col;
    __syncthreads();
    double pzv = 0.0;
    if (g == 0 && j < P.n) {
        row = ((s_row[0][c] + s_row[1][c]) + (s_row[2][c] + s_row[3][c])) +
              ((s_row[4][c] + s_row[5][c]) + (s_row[6][c] + s_row[7][c]));
        col = ((s_col[0][c] + s_col[1][c]) + (s_col[2][c] + s_col[3][c])) +
              ((s_col[4][c] + s_col[5][c]) + (s_col[6][c] + s_col[7][c]));
        if (P.alpha) row *= P.alpha[j];
        if (P.beta) col *= P.beta[j];
        const double zj = row + col;
        P.z[j] = zj;
        if (PEER) {
            const long long slot = ((long long)P.X.parity * P.X.nranks + P.X.rank) * P.X.slotStride;
            for (int h = 0; h < P.X.nranks; ++h) P.X.slots[h][slot + j] = zj;
        }
        if (P.F.w) {
            const double pj = fma(*P.F.beta, P.F.p[j], P.F.w[j]);   // same arithmetic as vec_at
            P.F.p[j] = pj;
            pzv = pj * zj;
        }
    }
    if (P.F.pz_part) {   // uniform
        if (g == 0) warp_part<false>(pzv, s_dot, tid);
        __syncthreads();
        if (tid == 0) {
            P.F.pz_part[T] = join4<false>(s_dot);
            if (PEER) __threadfence_system();   // this CTA's peer stores before its ticket
            s_last = take_ticket(P.F.ticket, gridDim.x) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const double pz =
                final_reduce<false>(P.F.pz_part, (int)gridDim.x, s_fin, tid, kRedThreads);
            if (tid == 0) {
                *P.F.pz = pz;
                *P.F.ticket = 0;
                __threadfence();
            }
            if (PEER && tid < P.X.nranks) {
                // every CTA has published: hand p.z over and raise the flag on rank `tid`
                const long long slot =
                    ((long long)P.X.parity * P.X.nranks + P.X.rank) * P.X.slotStride;
                P.X.slots[tid][slot + P.n] = pz;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long*>(
                    P.X.flags[tid] + P.X.parity * P.X.nranks + P.X.rank) = P.X.seq;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_f64_to_f32(const double* __restrict__ a,
                                                    float* __restrict__ out, long long count) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)a[i];
}

}  // namespace

// cached task table / workspace of a context, rebuilt when the row blocks change
struct SymvPlan {
    long long n = -1;
    LowerLayout L{0, 0, 0, 0, 0, 0, 0};
    int nTasks = 0;
    int grid = 0;
    long long wcols = 0;
    SymvTask* d_tasks = nullptr;
    double* d_Wc = nullptr;
    double* d_Wr = nullptr;
    double* d_Wh = nullptr;
    int* d_headSlot = nullptr;
    int* d_headColBegin = nullptr;
    int* d_headColSlots = nullptr;
};

void symv_plan_free(icqb_ctx* ctx) {
    SymvPlan* p = ctx->symv;
    if (!p) return;
    dev_free(ctx, p->d_tasks);
    dev_free(ctx, p->d_Wc);
    dev_free(ctx, p->d_Wr);
    dev_free(ctx, p->d_Wh);
    dev_free(ctx, p->d_headSlot);
    dev_free(ctx, p->d_headColBegin);
    dev_free(ctx, p->d_headColSlots);
    delete p;
    ctx->symv = nullptr;
}

static int symv_plan(icqb_ctx* ctx) {
    SymvPlan* p = ctx->symv;
    if (p && p->n == ctx->ntri && p->L.r0 == ctx->r0 && p->L.r1 == ctx->r1 &&
        p->L.q0 == ctx->q0 && p->L.q1 == ctx->q1)
        return ICQB_OK;
    symv_plan_free(ctx);
    LowerLayout L;
    int st = lower_layout(ctx, &L);
    if (st != ICQB_OK) return st;
    p = new SymvPlan();
    ctx->symv = p;
    p->n = ctx->ntri;
    p->L = L;
    p->wcols = (p->n + kT - 1) / kT * kT;
    p->grid = ctx->num_sms > 0 ? ctx->num_sms : 148;
    std::vector<SymvTask> tasks;
    for (int lt = 0; lt < L.nLocal(); ++lt) {
        const int I = (int)L.tileRow(lt);
        for (int J = 0; J <= I; ++J) tasks.push_back(SymvTask{lt, J});
    }
    p->nTasks = (int)tasks.size();
    if ((long long)p->nTasks != L.numTiles())
        return set_error(ctx, ICQB_ESTATE, "symv: tile enumeration does not match the layout");
    if (p->nTasks == 0) return ICQB_OK;

    // the static partition of the chunk sequence over the warps (same arithmetic as the
    // kernel): a span that starts inside an off-diagonal tile owns a head slot
    const long long nW = (long long)p->grid * kWarps;
    const long long total = (long long)p->nTasks * kChunksPerTile;
    const int nCols = (int)((p->n + kT - 1) / kT);
    std::vector<int> headSlot((size_t)nW, -1), slotCol;
    for (long long gw = 0; gw < nW; ++gw) {
        const long long c0 = total * gw / nW, c1 = total * (gw + 1) / nW;
        if (c1 <= c0 || c0 % kChunksPerTile == 0) continue;
        const SymvTask& t = tasks[(size_t)(c0 / kChunksPerTile)];
        if (t.J == (int)L.tileRow(t.lt)) continue;   // diagonal tile: no column part
        headSlot[(size_t)gw] = (int)slotCol.size();
        slotCol.push_back(t.J);
    }
    const int slots = (int)slotCol.size();
    // CSR of the head slots by column tile (counting sort: slots stay in ascending order)
    std::vector<int> headColBegin((size_t)nCols + 1, 0), headColSlots((size_t)std::max(slots, 1), 0);
    for (int sidx = 0; sidx < slots; ++sidx) headColBegin[(size_t)slotCol[sidx] + 1] += 1;
    for (int J = 0; J < nCols; ++J) headColBegin[(size_t)J + 1] += headColBegin[(size_t)J];
    {
        std::vector<int> fill(headColBegin.begin(), headColBegin.end() - 1);
        for (int sidx = 0; sidx < slots; ++sidx) headColSlots[(size_t)fill[(size_t)slotCol[sidx]]++] = sidx;
    }
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_tasks, sizeof(SymvTask) * tasks.size()));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_Wc, sizeof(double) * (size_t)L.nLocal() * p->wcols));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_Wr, sizeof(double) * (size_t)p->nTasks * kT));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_Wh, sizeof(double) * (size_t)std::max(slots, 1) * kT));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_headSlot, sizeof(int) * headSlot.size()));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_headColBegin, sizeof(int) * headColBegin.size()));
    ICQB_CUDA(ctx, dev_alloc(ctx, &p->d_headColSlots, sizeof(int) * headColSlots.size()));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_tasks, tasks.data(), sizeof(SymvTask) * tasks.size(),
                              cudaMemcpyHostToDevice));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_headSlot, headSlot.data(), sizeof(int) * headSlot.size(),
                              cudaMemcpyHostToDevice));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_headColBegin, headColBegin.data(),
                              sizeof(int) * headColBegin.size(), cudaMemcpyHostToDevice));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_headColSlots, headColSlots.data(),
                              sizeof(int) * headColSlots.size(), cudaMemcpyHostToDevice));
    // (every stored off-diagonal tile is begun by exactly one warp, which always writes the
    // tile's Wc slot, so Wc needs no initialisation)
    ICQB_CUDA(ctx, cudaFuncSetAttribute(k_symv, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes));
    ICQB_CUDA(ctx, cudaFuncSetAttribute(k_symv_f32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSmemBytes32));
    return ICQB_OK;
}

int launch_symv(icqb_ctx* ctx, const double* dA, const double* dx, const double* dgamma,
                const double* dalpha, const double* dbeta, double* dz, cudaStream_t stream,
                const int* skip, const SymvFuse* fuse, const SymvPeer* peer) {
    if (ctx->ntri == 0) return ICQB_OK;
    if ((reinterpret_cast<uintptr_t>(dA) & 15) != 0)
        return set_error(ctx, ICQB_EARG, "symv: matrix must be 16-byte aligned");
    int st = symv_plan(ctx);
    if (st != ICQB_OK) return st;
    SymvPlan* p = ctx->symv;
    SymvParams P;
    P.A = dA;
    P.n = p->n;
    P.L = p->L;
    P.nTasks = p->nTasks;
    P.tasks = p->d_tasks;
    P.x = dx;
    P.gamma = dgamma;
    P.alpha = dalpha;
    P.beta = dbeta;
    P.Wc = p->d_Wc;
    P.Wr = p->d_Wr;
    P.Wh = p->d_Wh;
    P.headSlot = p->d_headSlot;
    P.headColBegin = p->d_headColBegin;
    P.headColSlots = p->d_headColSlots;
    P.wcols = p->wcols;
    P.z = dz;
    P.skip = skip;
    if (fuse) P.F = *fuse;
    if (peer) {
        if (!fuse || !fuse->pz_part)
            return set_error(ctx, ICQB_EARG, "symv: the peer exchange needs the fused p.z");
        P.X = *peer;
    }
    if (p->nTasks > 0) {
        k_symv<<<p->grid, kThreads, kSmemBytes, stream>>>(P);
        ctx->launches += 1;
    }
    if (peer)
        k_symv_reduce<true><<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, kRedGroups), 0, stream>>>(P);
    else
        k_symv_reduce<false><<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, kRedGroups), 0, stream>>>(P);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_symv launch");
}

// Optional: one product of the mixed-precision CG.  Both matrix kernels are enqueued; the
// device flags decide which one does the work (skip64 / skip32 are maintained by k_cg_close:
// exactly one of them is 0 while the iteration runs), the reduce kernel follows either.
int launch_symv_mixed(icqb_ctx* ctx, const double* dA, const float* dA32, const double* dx,
                      const double* dgamma, const double* dalpha, const double* dbeta, double* dz,
                      cudaStream_t stream, const int* done, const int* skip64, const int* skip32,
                      const SymvFuse* fuse, const SymvPeer* peer) {
    if (ctx->ntri == 0) return ICQB_OK;
    if ((reinterpret_cast<uintptr_t>(dA) & 15) != 0 || (reinterpret_cast<uintptr_t>(dA32) & 15) != 0)
        return set_error(ctx, ICQB_EARG, "symv: matrices must be 16-byte aligned");
    int st = symv_plan(ctx);
    if (st != ICQB_OK) return st;
    SymvPlan* p = ctx->symv;
    SymvParams P;
    P.A = dA;
    P.A32 = dA32;
    P.n = p->n;
    P.L = p->L;
    P.nTasks = p->nTasks;
    P.tasks = p->d_tasks;
    P.x = dx;
    P.gamma = dgamma;
    P.alpha = dalpha;
    P.beta = dbeta;
    P.Wc = p->d_Wc;
    P.Wr = p->d_Wr;
    P.Wh = p->d_Wh;
    P.headSlot = p->d_headSlot;
    P.headColBegin = p->d_headColBegin;
    P.headColSlots = p->d_headColSlots;
    P.wcols = p->wcols;
    P.z = dz;
    if (fuse) P.F = *fuse;
    if (peer) {
        if (!fuse || !fuse->pz_part)
            return set_error(ctx, ICQB_EARG, "symv: the peer exchange needs the fused p.z");
        P.X = *peer;
    }
    if (p->nTasks > 0) {
        P.skip = skip64;
        k_symv<<<p->grid, kThreads, kSmemBytes, stream>>>(P);
        P.skip = skip32;
        k_symv_f32<<<p->grid, kThreads, kSmemBytes32, stream>>>(P);
        ctx->launches += 2;
    }
    P.skip = done;
    if (peer)
        k_symv_reduce<true><<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, kRedGroups), 0, stream>>>(P);
    else
        k_symv_reduce<false><<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, kRedGroups), 0, stream>>>(P);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_symv (mixed) launch");
}

// float32 copy of the locally stored tiles (count = stored tiles * 128 * 128)
int launch_f64_to_f32(icqb_ctx* ctx, const double* dA, float* dA32, long long count,
                      cudaStream_t stream) {
    if (count <= 0) return ICQB_OK;
    k_f64_to_f32<<<148 * 8, 256, 0, stream>>>(dA, dA32, count);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_f64_to_f32 launch");
}

// plain float32 product (tests): z = G x with the float32 copy
int launch_symv_f32_only(icqb_ctx* ctx, const float* dA32, const double* dx, const double* dgamma,
                         const double* dalpha, const double* dbeta, double* dz, cudaStream_t stream) {
    if (ctx->ntri == 0) return ICQB_OK;
    int st = symv_plan(ctx);
    if (st != ICQB_OK) return st;
    SymvPlan* p = ctx->symv;
    SymvParams P;
    P.A = nullptr;
    P.A32 = dA32;
    P.n = p->n;
    P.L = p->L;
    P.nTasks = p->nTasks;
    P.tasks = p->d_tasks;
    P.x = dx;
    P.gamma = dgamma;
    P.alpha = dalpha;
    P.beta = dbeta;
    P.Wc = p->d_Wc;
    P.Wr = p->d_Wr;
    P.Wh = p->d_Wh;
    P.headSlot = p->d_headSlot;
    P.headColBegin = p->d_headColBegin;
    P.headColSlots = p->d_headColSlots;
    P.wcols = p->wcols;
    P.z = dz;
    P.skip = nullptr;
    if (p->nTasks > 0) {
        k_symv_f32<<<p->grid, kThreads, kSmemBytes32, stream>>>(P);
        ctx->launches += 1;
    }
    k_symv_reduce<false><<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, kRedGroups), 0, stream>>>(P);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_symv_f32 launch");
}

}  // namespace icqb
