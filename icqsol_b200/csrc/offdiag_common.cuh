// Shared pieces of the off-diagonal assembly kernels (assemble_offdiag.cu: the validated
// kernel; assemble_offdiag_ff.cu: the far-field split variant): launch parameters, the
// inline-PTX helpers (mbarrier, bulk copy, reciprocal square root) and the tile decoding.
// Everything sits in an anonymous namespace: each translation unit gets its own copy.
#pragma once

#include <math.h>

#include "icqb_internal.cuh"
#include "quad_tables.cuh"

namespace icqb {

struct OffDiagParams {
    const double* qp_soa;   // [48][npad]
    const double* qp_aos;   // [npad+128][48]
    const double* area2;    // [npad+128]
    const double* nrm;      // [npad+128][4]  (nx, ny, nz, unused)
    const double* verts;    // [ntri][9]
    double* G;
    double* K;
    double* Ku;             // lower mode only: Ku[i,j] = K[j,i] at the position of (i,j)
    long long ld;
    long long n, npad;
    long long r0, r1;       // full mode: the row block
    long long nObsTiles, nLeft, nRight;  // full mode tile counts: local rows, columns < r0, >= r1
    LowerLayout L;          // lower mode: row blocks and tile-contiguous layout
    int lower;              // 1 = block-lower-triangular storage
    // far-field split variant only (assemble_offdiag_ff.cu); unused by k_offdiag
    const double* bound = nullptr;      // [npad+128][4]  centroid, radius (max vertex distance)
    double kappa = 0.0;                 // a pair is far when |c_i - c_j| >= kappa (rho_i + rho_j)
    double guard2 = 0.0;                // ... and |c_i - c_j|^2 >= guard2
    unsigned long long* stats = nullptr;   // [2] (warp, source triangle) pairs: fast, all
};

// far-field split variant (assemble_offdiag_ff.cu): fills bound/kappa/guard2/stats and launches
int launch_offdiag_ff(icqb_ctx* ctx, const OffDiagParams& params, dim3 grid, bool withK);

namespace {

constexpr int kTileObs = 128;   // threads per CTA = observer triangles per CTA
constexpr int kTileSrc = 128;   // source triangles per CTA
constexpr int kSub = 32;        // source triangles per bulk copy
constexpr int kSubBytes = kSub * kQpStride * 8;
constexpr int kStages = 2;


#ifdef ICQB_EMU   // CPU emulation build of the test suite (tools/emu): no PTX
// same refinement on a seed of single precision (the hardware seed has >= 22 bits)
__device__ __forceinline__ double rsqrt_1ulp(double x) {
    double y0 = (double)(float)(1.0 / sqrt(x));
    double t = x * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    double q = e * y0;
    return fma(p, q, y0);
}
// mbarrier modelled as (pending transaction bytes << 32) | completed phases; a bulk copy
// completes when issued, and the phase completes with the last expected byte
__device__ __forceinline__ void mbar_init(uint64_t* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    *bar += (uint64_t)bytes << 32;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((*reinterpret_cast<volatile uint64_t*>(bar) & 1) == parity) emu::yield();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    memcpy(dst, src, bytes);
    *bar -= (uint64_t)bytes << 32;
    if ((*bar >> 32) == 0) *bar += 1;
}
#else
__device__ __forceinline__ double rsqrt_1ulp(double x) {
    // MUFU.RSQ64H seed (rel. error <= 2^-22.4) + one third-order step:
    // 1/sqrt(x) = y0 (1 - e)^(-1/2), e = 1 - x y0^2  ->  y0 (1 + e/2 + 3 e^2/8),
    // truncation 5 e^3/16 < 2^-65.  No slow path: r^2 of two distinct triangles is
    // a normal number; for coincident points (r = 0) the result is inf/nan, as in
    // the reference (1.0/sqrt(0)).
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double t = x * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    double q = e * y0;
    return fma(p, q, y0);
}

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

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#endif

// Decode the CTA's tile pair.  Task order: the locally owned square block first
// (lower-triangular tile pairs incl. the diagonal tiles), then the column ranges
// left and right of it.
struct Tile {
    long long obs0;   // first observer triangle (global index)
    long long obsEnd; // one past the last observer triangle of the row block
    long long lrow0;  // full mode: local (storage) row of obs0
    long long tile;   // lower mode: index of this tile in the local tile sequence
    long long src0;   // first source triangle
    long long srcEnd; // one past the last valid source triangle of this tile's range
    int kind;         // 0 = diagonal tile, 1 = mirrored (I > J inside the block), 2 = direct only,
                      // 3 = lower storage, strictly below the diagonal tile
                      // (-1 = nothing to do: no longer produced, the grids hold stored tiles only)
};

__device__ __forceinline__ Tile decode_tile(const OffDiagParams& P, long long task) {
    if (P.lower) {
        // 1-D grid over the STORED tiles only (J <= I): task = position of the tile in the local
        // tile sequence.  Strip lt of a row block whose first tile row is I0 starts at
        // k (I0 + 1) + k (k - 1) / 2, k = the strip's index inside its block.
        const bool second = task >= P.L.tiles0;
        const long long t = second ? task - P.L.tiles0 : task;
        const long long I0 = (second ? P.L.q0 : P.L.r0) / kTileObs;
        const double bq = 2.0 * (double)I0 + 1.0;
        long long k = (long long)((sqrt(bq * bq + 8.0 * (double)t) - bq) * 0.5);
        if (k < 0) k = 0;
        while (k * (I0 + 1) + k * (k - 1) / 2 > t) --k;
        while ((k + 1) * (I0 + 1) + (k + 1) * k / 2 <= t) ++k;
        const int lt = (int)k + (second ? P.L.nTiles0 : 0);
        const long long J = t - (k * (I0 + 1) + k * (k - 1) / 2);
        Tile r;
        r.obs0 = P.L.obs0(lt);
        r.obsEnd = P.L.obsEnd(lt);
        r.lrow0 = 0;
        const long long I = r.obs0 / kTileObs;
        r.tile = task;
        r.src0 = J * kTileSrc;
        r.srcEnd = (J == I) ? min(P.n, r.src0 + kTileSrc) : r.src0 + kTileSrc;
        r.kind = (J == I) ? 0 : 3;
        return r;
    }
    Tile t;
    const long long nT = P.nObsTiles;
    const long long nTri = nT * (nT + 1) / 2;
    if (task < nTri) {
        // task = I(I+1)/2 + J, 0 <= J <= I
        long long I = (long long)((sqrt(8.0 * (double)task + 1.0) - 1.0) * 0.5);
        while (I * (I + 1) / 2 > task) --I;
        while ((I + 1) * (I + 2) / 2 <= task) ++I;
        long long J = task - I * (I + 1) / 2;
        t.obs0 = P.r0 + I * kTileObs;
        t.obsEnd = P.r1;
        t.lrow0 = I * kTileObs;
        t.tile = 0;
        t.src0 = P.r0 + J * kTileSrc;
        t.srcEnd = P.r1;
        t.kind = (I == J) ? 0 : 1;
        return t;
    }
    task -= nTri;
    const long long nSide = P.nLeft + P.nRight;
    long long I = task / nSide, J = task - I * nSide;
    t.obs0 = P.r0 + I * kTileObs;
    t.obsEnd = P.r1;
    t.lrow0 = I * kTileObs;
    t.tile = 0;
    t.kind = 2;
    if (J < P.nLeft) {
        t.src0 = J * kTileSrc;
        t.srcEnd = P.r0;
    } else {
        t.src0 = P.r1 + (J - P.nLeft) * kTileSrc;
        t.srcEnd = P.n;
    }
    return t;
}

}  // namespace

}  // namespace icqb
