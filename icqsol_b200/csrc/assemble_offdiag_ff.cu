// Far-field split of the fused G/K off-diagonal assembly (assembly variant 1,
// icqb_set_assembly_variant).  First run on a B200 in round 2 (gpurun_out/r2a_*, r2b_*): every
// G entry within 5e-15 of the direct kernel's; the direct kernel (variant 0, assemble_offdiag.cu)
// stays the default until this one is faster by more than a few per cent.
//
// Same CTA shape, tile order and stores as k_offdiag.  What changes:
//
//   * The arithmetic of one kernel evaluation for pairs of triangles that are well separated.
//     The observer points of a flat triangle are affine in the rule's nodes:
//     o_k = a + xi_k (b - a) + eta_k (c - a).  With the origin moved to the observer's vertex
//     a (s' = s - a, u_k = o_k - a),
//         r^2 = |s' - u_k|^2 = (|s'|^2 + |u_k|^2) + xi_k (-2 s'.(b-a)) + eta_k (-2 s'.(c-a)),
//     so per SOURCE POINT the thread forms A = |s'|^2, B = -2 s'.(b-a), C = -2 s'.(c-a)
//     (12 FP64 instructions, amortised over 16 evaluations) and per evaluation
//         r2 = fma(xi_k, B, fma(eta_k, C, A + q_k)),   q_k = |u_k|^2 kept in registers:
//     1 DADD + 2 DFMA instead of 3 DADD + DMUL + 2 DFMA, and 16 registers of q instead of 48
//     of observer points.
//   * Moving the origin to the observer triangle keeps the cancellation harmless only when
//     the two triangles are far apart compared with their size: the three terms are then all
//     of the size of r^2 (relative rounding error of r^2 <= ~16 eps).
//     Every lane tests |c_i - c_j| >= kappa (rho_i + rho_j) for its pair of triangles
//     (centroids c, radii rho = largest centroid-vertex distance, kappa = 3) and the warp
//     votes: only if all 32 pairs are far does the warp take the fast arithmetic for this
//     source triangle; otherwise it evaluates the 256 points exactly as k_offdiag does, from
//     the reference's rounded quadrature points (bem/icqQuadrature.cpp:233-239), which the
//     thread re-forms from its own vertices with the rounding order of k_prepare.  A second
//     condition, |c_i - c_j| >= max|coordinate| / 512, bounds the difference between the
//     affine points used here and the reference's rounded ones (<= 1.5 eps max|coordinate|
//     each) to <= 2e-13 of r.
//   * K: both triangles are flat, so n_j.(x - y) depends on the observer point only and is
//     affine in the nodes as well; the mirrored n_i.(y - x) is n_i.s'.
//   * The warps of a CTA no longer run in lockstep (one takes the near path while the others
//     do not), so the 2-stage ring with a __syncthreads per refill of k_offdiag would stall
//     them (ncu, round 2: 4.8 % of the warp samples at that barrier).  Here the whole source
//     tile -- 128 triangles: quadrature points, bounds + first vertex, normals, 60 KB -- is
//     requested up front as four sub-tiles, each on its own mbarrier, and no warp ever waits
//     for another.
//
// FP64-pipe instructions per evaluation on the fast path (SASS counts in DESIGN.md 4.1b):
// G 9 instead of 12, G+K 13 instead of 16, plus the per-source-point work above.
#include <stdlib.h>

#include "offdiag_common.cuh"

namespace icqb {

namespace {

// The order-8 rule as literals (bem/icqQuadrature.cpp:122-144, the 14-digit values of
// quad_tables.cuh): identical literals share one constant, so the 48 numbers of the rule
// need 15 uniform registers' worth of distinct values (10 coordinates, 5 weights) instead of
// 48 loads from c_triNodes -- the unrolled loops below would otherwise spill uniform registers.
#define ICQB_TRI8(X)                                              \
    X(0, 0.33333333333333, 0.33333333333333, 0.14431560767779)    \
    X(1, 0.45929258829272, 0.45929258829272, 0.09509163426728)    \
    X(2, 0.45929258829272, 0.08141482341455, 0.09509163426728)    \
    X(3, 0.08141482341455, 0.45929258829272, 0.09509163426728)    \
    X(4, 0.17056930775176, 0.17056930775176, 0.10321737053472)    \
    X(5, 0.17056930775176, 0.65886138449648, 0.10321737053472)    \
    X(6, 0.65886138449648, 0.17056930775176, 0.10321737053472)    \
    X(7, 0.05054722831703, 0.05054722831703, 0.03245849762320)    \
    X(8, 0.05054722831703, 0.89890554336594, 0.03245849762320)    \
    X(9, 0.89890554336594, 0.05054722831703, 0.03245849762320)    \
    X(10, 0.26311282963464, 0.72849239295540, 0.02723031417443)   \
    X(11, 0.72849239295540, 0.00839477740996, 0.02723031417443)   \
    X(12, 0.00839477740996, 0.26311282963464, 0.02723031417443)   \
    X(13, 0.72849239295540, 0.26311282963464, 0.02723031417443)   \
    X(14, 0.26311282963464, 0.00839477740996, 0.02723031417443)   \
    X(15, 0.00839477740996, 0.72849239295540, 0.02723031417443)
#define ICQB_CHECK_NODE(K, XI, ETA, W)                                                       \
    static_assert(kTriNodes[46 + K].xi == XI && kTriNodes[46 + K].eta == ETA &&              \
                      kTriNodes[46 + K].w == W,                                               \
                  "ICQB_TRI8 must repeat the order-8 rows of quad_tables.cuh");
ICQB_TRI8(ICQB_CHECK_NODE)
#undef ICQB_CHECK_NODE
static_assert(kNq == 16, "ICQB_TRI8 lists 16 nodes");

// state of one observer triangle (one thread)
struct Obs {
    double ax, ay, az;       // vertex a
    double b2x, b2y, b2z;    // -2 (b - a)
    double c2x, c2y, c2z;    // -2 (c - a)
    double q[kNq];           // |o_k - a|^2 of the affine points
    double nx, ny, nz;       // unit normal (K)
    double cx, cy, cz, rho;  // centroid, radius
};

struct PairSums {
    double g;    // sum_i0 sum_i1 w_i0 w_i1 / r
    double k;    // sum w w n_j.(x - y) / r^3   -> K[i,j]
    double km;   // sum w w n_i.(y - x) / r^3   -> K[j,i]
};

// per-triangle record next to the quadrature points of a source sub-tile (k_bounds):
// centroid, radius, first vertex
constexpr int kBndStride = 8;
// the whole source tile in shared memory: kFfSub sub-tiles of 32 triangles, one mbarrier each
constexpr int kFfSub = kTileSrc / kSub;
constexpr int kFfSrcDoubles = kSub * kQpStride;      // 1536
constexpr int kFfBndDoubles = kSub * kBndStride;     // 256
constexpr int kFfNrmDoubles = kSub * 4;              // 128
constexpr int kFfStageDoubles = kFfSrcDoubles + kFfBndDoubles + kFfNrmDoubles;
// Row-coalesced stores: a lane owns a ROW of the tile, so storing each entry as it is produced
// writes 8 bytes at a 1 KB stride -- partial sectors that the L2 has to read-fill and write
// twice (ncu, round 1: 6.41 GB of DRAM traffic for 4.82 GB of matrices).  Instead every warp
// parks the entries of kStCols consecutive source triangles in its own shared-memory patch
// ([matrix][32 rows][kStCols + 1]) and then writes, per matrix, 32 row segments of
// kStCols * 8 = 64 contiguous bytes (two complete sectors each).
constexpr int kStCols = 8;
constexpr int kStPitch = kStCols + 1;
constexpr int kStWarpDoubles = 3 * 32 * kStPitch;                  // G, K, Ku
constexpr int kFfBarBytes = 64;
constexpr int kFfSmemBytes = kFfSub * kFfStageDoubles * 8 + kFfBarBytes + (kTileObs / 32) * kStWarpDoubles * 8;   // 89,216 B

// well separated pair: affine observer points, origin at the observer's vertex a.
// Software-pipelined: the per-source-point work of point i1 + 1 (three shared-memory loads, then
// the short dependent chains that form A, B, C and n_i . s') is issued inside the evaluation
// block of point i1, where the scheduler can slot it between the sixteen independent
// evaluations -- with two warps per scheduler (255 registers) nothing else hides that latency
// at the head of every iteration (B200, config C2, G+K: 50.6 -> 47.9 ms, identical entries).  The
// loop over the source points is unrolled twice: no gain for G+K, 1.8 % for G alone (fewer
// re-materialised constants).
template <bool WITH_K>
__device__ __forceinline__ PairSums pair_fast(const Obs& O, const double* __restrict__ pj,
                                              const double* __restrict__ nj,
                                              const double* __restrict__ bj) {
    PairSums S{0.0, 0.0, 0.0};
    double r3row[WITH_K ? kNq : 1];
    if (WITH_K) {
#pragma unroll
        for (int k = 0; k < kNq; ++k) r3row[k] = 0.0;
    }
    double nA = 0.0, nB = 0.0, nC = 0.0, nNd = 0.0;
    auto point = [&](int i1, double& A, double& B, double& C, double& nd) {
        const double sx = pj[3 * i1 + 0] - O.ax;
        const double sy = pj[3 * i1 + 1] - O.ay;
        const double sz = pj[3 * i1 + 2] - O.az;
        A = fma(sz, sz, fma(sy, sy, sx * sx));
        B = fma(sz, O.b2z, fma(sy, O.b2y, sx * O.b2x));
        C = fma(sz, O.c2z, fma(sy, O.c2y, sx * O.c2x));
        // mirrored entry K[j,i]: n_i . (y - a_i) = n_i . s'
        if (WITH_K) nd = fma(O.nz, sz, fma(O.ny, sy, O.nx * sx));
    };
    point(0, nA, nB, nC, nNd);
#pragma unroll 2
    for (int i1 = 0; i1 < kNq; ++i1) {
        const double A = nA, B = nB, C = nC, nd = nNd;
        // (the last iteration re-forms point 15: one wasted set-up in sixteen, no branch)
        point(i1 + 1 < kNq ? i1 + 1 : kNq - 1, nA, nB, nC, nNd);
        const double w1 = c_triNodes[kTri8 + i1].w;
        double in[2] = {0.0, 0.0};   // two chains of eight: enough with sixteen evaluations in flight
        double c3[2] = {0.0, 0.0};
#define ICQB_FF_EVAL(K, XI, ETA, W)                                            \
    {                                                                          \
        /* fma(XI, B, A) is shared by the nodes with the same xi (10 distinct values) */ \
        const double r2 = fma(ETA, C, fma(XI, B, A) + O.q[K]);                 \
        const double ri = rsqrt_1ulp(r2);                                      \
        in[K & 1] = fma(W, ri, in[K & 1]);                                     \
        if (WITH_K) {                                                          \
            const double r3 = ri * ri * ri;                                    \
            r3row[K] = fma(w1, r3, r3row[K]);                                  \
            c3[K & 1] = fma(W, r3, c3[K & 1]);                                 \
        }                                                                      \
    }
        ICQB_TRI8(ICQB_FF_EVAL)
#undef ICQB_FF_EVAL
        const double in0 = in[0], in1 = in[1];
        const double c30 = c3[0], c31 = c3[1];
        S.g = fma(w1, in0 + in1, S.g);
        if (WITH_K) S.km = fma(w1 * nd, c30 + c31, S.km);
    }
    if (WITH_K) {
        // K[i,j]: n_j . (o_k - a_j) = n_j.(a_i - a_j) - (xi_k n_j.b2 + eta_k n_j.c2) / 2
        const double njx = nj[0], njy = nj[1], njz = nj[2];
        const double e0 = fma(njz, O.az - bj[6], fma(njy, O.ay - bj[5], njx * (O.ax - bj[4])));
        const double nb = fma(njz, O.b2z, fma(njy, O.b2y, njx * O.b2x));
        const double nc = fma(njz, O.c2z, fma(njy, O.c2y, njx * O.c2x));
        // sum_k w_k (e0 - (xi_k nb + eta_k nc) / 2) R_k as three moment sums of the row sums R_k
        // with literal weights (w_k, w_k xi_k, w_k eta_k): 3 DFMA per node instead of 5 instructions
        double m0[2] = {0.0, 0.0}, m1[2] = {0.0, 0.0}, m2[2] = {0.0, 0.0};
#define ICQB_FF_K(K, XI, ETA, W)                                   \
    {                                                              \
        m0[K & 1] = fma(W, r3row[K], m0[K & 1]);                   \
        m1[K & 1] = fma(W * XI, r3row[K], m1[K & 1]);              \
        m2[K & 1] = fma(W * ETA, r3row[K], m2[K & 1]);             \
    }
        ICQB_TRI8(ICQB_FF_K)
#undef ICQB_FF_K
        S.k = fma(e0, m0[0] + m0[1], -0.5 * fma(nb, m1[0] + m1[1], nc * (m2[0] + m2[1])));
    }
    return S;
}

// near pair: the arithmetic of k_offdiag on the reference's rounded quadrature points
// p = pa + xi db + eta dc (bem/icqQuadrature.cpp:233-234), re-formed from the thread's own
// vertices with the explicitly rounded operations of k_prepare: bit-identical to the tables
template <bool WITH_K>
__device__ __forceinline__ PairSums pair_exact(const Obs& O, const double* __restrict__ pj,
                                               const double* __restrict__ nj,
                                               const double* __restrict__ bj) {
    PairSums S{0.0, 0.0, 0.0};
    double c3[WITH_K ? kNq : 1];
    if (WITH_K) {
#pragma unroll
        for (int k = 0; k < kNq; ++k) c3[k] = 0.0;
    }
    double njx = 0, njy = 0, njz = 0, pjax = 0, pjay = 0, pjaz = 0;
    if (WITH_K) {
        njx = nj[0], njy = nj[1], njz = nj[2];
        pjax = bj[4], pjay = bj[5], pjaz = bj[6];
    }
    // b - a and c - a: O.b2 = -2 (b - a) exactly, so -0.5 * O.b2 is the rounded difference
    const double dbx = -0.5 * O.b2x, dby = -0.5 * O.b2y, dbz = -0.5 * O.b2z;
    const double dcx = -0.5 * O.c2x, dcy = -0.5 * O.c2y, dcz = -0.5 * O.c2z;
#pragma unroll 1
    for (int i0 = 0; i0 < kNq; ++i0) {
        const double xi = c_triNodes[kTri8 + i0].xi, eta = c_triNodes[kTri8 + i0].eta;
        const double ox = __dadd_rn(__dadd_rn(O.ax, __dmul_rn(xi, dbx)), __dmul_rn(eta, dcx));
        const double oy = __dadd_rn(__dadd_rn(O.ay, __dmul_rn(xi, dby)), __dmul_rn(eta, dcy));
        const double oz = __dadd_rn(__dadd_rn(O.az, __dmul_rn(xi, dbz)), __dmul_rn(eta, dcz));
        const double w0 = c_triNodes[kTri8 + i0].w;
        double in0 = 0.0, in1 = 0.0, in2 = 0.0, in3 = 0.0;
        double r30 = 0.0, r31 = 0.0;
#pragma unroll
        for (int i1 = 0; i1 < kNq; ++i1) {
            const double dx = pj[3 * i1 + 0] - ox;
            const double dy = pj[3 * i1 + 1] - oy;
            const double dz = pj[3 * i1 + 2] - oz;
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const double ri = rsqrt_1ulp(r2);
            const double w1 = c_triNodes[kTri8 + i1].w;
            if ((i1 & 3) == 0) in0 = fma(w1, ri, in0);
            if ((i1 & 3) == 1) in1 = fma(w1, ri, in1);
            if ((i1 & 3) == 2) in2 = fma(w1, ri, in2);
            if ((i1 & 3) == 3) in3 = fma(w1, ri, in3);
            if (WITH_K) {
                const double r3 = ri * ri * ri;
                c3[i1] = fma(w0, r3, c3[i1]);
                if (i1 & 1) r31 = fma(w1, r3, r31);
                else r30 = fma(w1, r3, r30);
            }
        }
        S.g = fma(w0, (in0 + in1) + (in2 + in3), S.g);
        if (WITH_K) {
            // K[i,j]: n_j . (x - y) = n_j . (o - a_j)
            const double nd = fma(njz, oz - pjaz, fma(njy, oy - pjay, njx * (ox - pjax)));
            S.k = fma(w0 * nd, r30 + r31, S.k);
        }
    }
    if (WITH_K) {
        // K[j,i]: n_i . (y - x) = n_i . (s - a_i)
#pragma unroll
        for (int i1 = 0; i1 < kNq; ++i1) {
            const double nd = fma(O.nz, pj[3 * i1 + 2] - O.az,
                                  fma(O.ny, pj[3 * i1 + 1] - O.ay, O.nx * (pj[3 * i1 + 0] - O.ax)));
            S.km = fma(c_triNodes[kTri8 + i1].w * nd, c3[i1], S.km);
        }
    }
    return S;
}

template <bool WITH_K>
__global__ void __launch_bounds__(kTileObs) k_offdiag_ff(const OffDiagParams P) {
    // [kFfSub] x { quadrature points 32 x 48 | bounds 32 x 8 | normals 32 x 4 }, then the barriers
    extern __shared__ __align__(128) double s_tile[];
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tile + kFfSub * kFfStageDoubles);
    double* s_park = s_tile + kFfSub * kFfStageDoubles + kFfBarBytes / 8 + (threadIdx.x >> 5) * kStWarpDoubles;

    const int tid = threadIdx.x;
    const Tile T = decode_tile(P, blockIdx.x);
    if (T.kind < 0) return;  // lower storage: tile above the diagonal (whole CTA exits)

    const long long srcLeft = T.srcEnd - T.src0;
    const int nSubTiles =
        (int)(srcLeft >= kTileSrc ? kTileSrc / kSub : (srcLeft + kSub - 1) / kSub);
    if (tid == 0) {
        for (int s = 0; s < kFfSub; ++s) mbar_init(&s_bar[s], 1);
#ifndef ICQB_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        // the whole source tile at once: the tables are padded by 128 rows beyond n, so that
        // complete sub-tiles can always be copied
        for (int s = 0; s < nSubTiles; ++s) {
            double* dst = s_tile + s * kFfStageDoubles;
            const long long j0 = T.src0 + (long long)s * kSub;
            mbar_expect_tx(&s_bar[s], (kFfSrcDoubles + kFfBndDoubles + (WITH_K ? kFfNrmDoubles : 0)) * 8);
            bulk_g2s(dst, P.qp_aos + j0 * kQpStride, kFfSrcDoubles * 8, &s_bar[s]);
            bulk_g2s(dst + kFfSrcDoubles, P.bound + j0 * kBndStride, kFfBndDoubles * 8, &s_bar[s]);
            if (WITH_K)
                bulk_g2s(dst + kFfSrcDoubles + kFfBndDoubles, P.nrm + j0 * 4, kFfNrmDoubles * 8, &s_bar[s]);
        }
    }
    __syncthreads();   // the barriers are initialised before anybody waits on them

    // observer triangle of this thread
    const long long iObs = T.obs0 + tid;
    const bool obsValid = iObs < T.obsEnd;
    const long long iLoad = obsValid ? iObs : (T.obsEnd - 1);
    Obs O;
    {
        const double* v = P.verts + 9 * iLoad;
        O.ax = __ldg(v + 0);
        O.ay = __ldg(v + 1);
        O.az = __ldg(v + 2);
        // b - a and c - a are exact or correctly rounded, as in k_prepare; the factor -2 is exact
        O.b2x = -2.0 * (__ldg(v + 3) - O.ax);
        O.b2y = -2.0 * (__ldg(v + 4) - O.ay);
        O.b2z = -2.0 * (__ldg(v + 5) - O.az);
        O.c2x = -2.0 * (__ldg(v + 6) - O.ax);
        O.c2y = -2.0 * (__ldg(v + 7) - O.ay);
        O.c2z = -2.0 * (__ldg(v + 8) - O.az);
#define ICQB_FF_Q(K, XI, ETA, W)                                                   \
    {                                                                              \
        const double ux = fma(XI, O.b2x, ETA * O.c2x); /* -2 (o_K - a) */          \
        const double uy = fma(XI, O.b2y, ETA * O.c2y);                             \
        const double uz = fma(XI, O.b2z, ETA * O.c2z);                             \
        O.q[K] = 0.25 * fma(uz, uz, fma(uy, uy, ux * ux));                         \
    }
        ICQB_TRI8(ICQB_FF_Q)
#undef ICQB_FF_Q
        O.nx = O.ny = O.nz = 0.0;
        if (WITH_K) {
            O.nx = __ldg(P.nrm + 4 * iLoad + 0);
            O.ny = __ldg(P.nrm + 4 * iLoad + 1);
            O.nz = __ldg(P.nrm + 4 * iLoad + 2);
        }
        O.cx = __ldg(P.bound + kBndStride * iLoad + 0);
        O.cy = __ldg(P.bound + kBndStride * iLoad + 1);
        O.cz = __ldg(P.bound + kBndStride * iLoad + 2);
        O.rho = __ldg(P.bound + kBndStride * iLoad + 3);
    }
    const double a2Obs = __ldg(P.area2 + iLoad);
    // -1/(4 pi) * 0.5  (icqLaplaceFunctor.cpp:9, icqQuadrature.cpp:245)
    const double cG = 0.5 / (-4.0 * 3.14159265358979323846);

    // storage row of local row 0 of this CTA (the warp adds its rows when it flushes its patch)
    const long long rowPitch = P.lower ? (long long)kLowerTile : P.ld;
    const long long rowBase = P.lower ? T.tile * kLowerTileElems - T.src0 : T.lrow0 * P.ld;
    const int lane = tid & 31, wrow0 = tid & ~31;

    // the patch of this warp -> global memory: `count` columns starting at source triangle jf0;
    // 4 rows x 8 columns per instruction, 64 contiguous bytes per row
    auto flush = [&](long long jf0, int count) {
        __syncwarp();
        const int col = lane & (kStCols - 1);
#pragma unroll
        for (int rr = 0; rr < 32; rr += 4) {
            const int row = rr + (lane >> 3);
            const long long iRow = T.obs0 + wrow0 + row;
            if (col < count && iRow < T.obsEnd) {
                const long long at = rowBase + (long long)(wrow0 + row) * rowPitch + jf0 + col;
                const bool offDiag = (T.kind != 0) || (jf0 + col != iRow);
                // (the diagonal of G belongs to k_diag: never written here)
                if (offDiag) P.G[at] = s_park[row * kStPitch + col];
                if (WITH_K) {
                    P.K[at] = s_park[(32 + row) * kStPitch + col];
                    if (P.lower) P.Ku[at] = s_park[(64 + row) * kStPitch + col];
                }
            }
        }
        __syncwarp();
    };

    unsigned int nFast = 0, nAll = 0;   // statistics (lane 0 of every warp adds them up)

    for (int sub = 0; sub < nSubTiles; ++sub) {
        const long long j0 = T.src0 + (long long)sub * kSub;
        mbar_wait(&s_bar[sub], 0);
        const double* sp = s_tile + sub * kFfStageDoubles;
        const double* sb = sp + kFfSrcDoubles;
        const double* sn = sb + kFfBndDoubles;

        for (int jj = 0; jj < kSub; ++jj) {
            const long long j = j0 + jj;
            if (j >= T.srcEnd) break;  // uniform
            const double* pj = sp + jj * kQpStride;
            const double* bj = sb + jj * kBndStride;
            const double* nj = sn + jj * 4;
            // separation test of this lane's pair, then the warp's vote
            const double ex = bj[0] - O.cx;
            const double ey = bj[1] - O.cy;
            const double ez = bj[2] - O.cz;
            const double d2 = fma(ez, ez, fma(ey, ey, ex * ex));
            const double lim = P.kappa * (O.rho + bj[3]);
            const bool farLane = (d2 >= lim * lim) && (d2 >= P.guard2);
            const bool far = __all_sync(0xffffffffu, farLane) != 0;
            nAll += 1;
            PairSums S;
            if (far) {
                nFast += 1;
                S = pair_fast<WITH_K>(O, pj, nj, bj);
            } else {
                S = pair_exact<WITH_K>(O, pj, nj, bj);
            }

            const bool offDiag = (T.kind != 0) || (j != iObs);
            const double a2Src = __ldg(P.area2 + j);
            const int pc = jj & (kStCols - 1);
            s_park[lane * kStPitch + pc] = (cG * a2Src) * S.g;
            if (WITH_K) {
                // K[i,i] = 0 (flat panel, principal value); K[j,i] is kept next to K[i,j]
                s_park[(32 + lane) * kStPitch + pc] = offDiag ? (cG * a2Src) * S.k : 0.0;
                s_park[(64 + lane) * kStPitch + pc] = offDiag ? (cG * a2Obs) * S.km : 0.0;
            }
            if (T.kind == 1 && obsValid) {
                // mirror: G[j,i] = g * areaObs / areaSrc  (icqLaplaceMatrices.cpp:98); consecutive
                // lanes write consecutive entries of row j
                P.G[(j - P.r0) * P.ld + iObs] = (cG * a2Obs) * S.g;
                if (WITH_K) P.K[(j - P.r0) * P.ld + iObs] = (cG * a2Obs) * S.km;
            }
            if (pc == kStCols - 1 || j + 1 >= T.srcEnd) flush(j - pc, pc + 1);
        }
    }
    if (P.stats && (tid & 31) == 0) {
        atomicAdd(P.stats + 0, (unsigned long long)nFast);
        atomicAdd(P.stats + 1, (unsigned long long)nAll);
    }
}

// centroid, radius (largest centroid-vertex distance, rounded up a little) and first vertex
// per triangle; the padding rows get a huge radius so that a padded triangle is never "far"
__global__ void __launch_bounds__(128) k_bounds(const double* __restrict__ verts, long long n,
                                                long long nalloc, double* __restrict__ bound) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nalloc) return;
    double* o = bound + kBndStride * t;
    if (t >= n) {
#pragma unroll
        for (int k = 0; k < kBndStride; ++k) o[k] = 0.0;
        o[3] = 1.0e300;
        return;
    }
    const double* v = verts + 9 * t;
    const double cx = (v[0] + v[3] + v[6]) * (1.0 / 3.0);
    const double cy = (v[1] + v[4] + v[7]) * (1.0 / 3.0);
    const double cz = (v[2] + v[5] + v[8]) * (1.0 / 3.0);
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double dx = v[3 * k + 0] - cx, dy = v[3 * k + 1] - cy, dz = v[3 * k + 2] - cz;
        r2 = fmax(r2, fma(dz, dz, fma(dy, dy, dx * dx)));
    }
    o[0] = cx;
    o[1] = cy;
    o[2] = cz;
    o[3] = sqrt(r2) * (1.0 + 1.0e-12);
    o[4] = v[0];
    o[5] = v[1];
    o[6] = v[2];
    o[7] = 0.0;
}

}  // namespace

constexpr double kFarKappa = 3.0;

// Prepare the per-triangle bounds (once per mesh) and launch the kernel.
int launch_offdiag_ff(icqb_ctx* ctx, const OffDiagParams& params, dim3 grid, bool withK) {
    OffDiagParams P = params;
    const long long nalloc = ctx->ntri_pad + kPad;
    if (!ctx->d_bound) {
        ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_bound, sizeof(double) * kBndStride * nalloc));
        k_bounds<<<(unsigned)((nalloc + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_verts, ctx->ntri,
                                                                            nalloc, ctx->d_bound);
        ctx->launches += 1;
        ICQB_CUDA(ctx, cudaGetLastError());
    }
    if (!ctx->d_ffstats) ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_ffstats, 2 * sizeof(unsigned long long)));
    ICQB_CUDA(ctx, cudaMemsetAsync(ctx->d_ffstats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    P.bound = ctx->d_bound;
    // ICQB_FF_KAPPA (experiments): separation factor of the far-field test, 2 <= kappa <= 16
    static const double kappa = [] {
        const char* e = getenv("ICQB_FF_KAPPA");
        const double v = e ? atof(e) : kFarKappa;
        return (v >= 2.0 && v <= 16.0) ? v : kFarKappa;
    }();
    P.kappa = kappa;
    const double guard = ctx->coord_max / 512.0;
    P.guard2 = guard * guard;
    P.stats = ctx->d_ffstats;
    static bool attr = false;
    if (!attr) {
        ICQB_CUDA(ctx, cudaFuncSetAttribute(k_offdiag_ff<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kFfSmemBytes));
        ICQB_CUDA(ctx, cudaFuncSetAttribute(k_offdiag_ff<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kFfSmemBytes));
        attr = true;
    }
    cudaEventRecord(ctx->ev0, ctx->stream);
    if (withK)
        k_offdiag_ff<true><<<grid, kTileObs, kFfSmemBytes, ctx->stream>>>(P);
    else
        k_offdiag_ff<false><<<grid, kTileObs, kFfSmemBytes, ctx->stream>>>(P);
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_offdiag_ff launch");
}

}  // namespace icqb

using namespace icqb;

extern "C" {

int icqb_set_assembly_variant(icqb_ctx* ctx, int variant) {
    if (!ctx) return ICQB_EARG;
    if (variant < 0 || variant > 1)
        return set_error(ctx, ICQB_EARG, "icqb_set_assembly_variant: variant must be 0 or 1");
    ctx->offdiag_variant = variant;
    return ICQB_OK;
}

int icqb_assembly_stats(icqb_ctx* ctx, int64_t* fast_pairs, int64_t* all_pairs) {
    if (!ctx) return ICQB_EARG;
    unsigned long long h[2] = {0, 0};
    if (ctx->d_ffstats) {
        ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
        ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ICQB_CUDA(ctx, cudaMemcpy(h, ctx->d_ffstats, sizeof(h), cudaMemcpyDeviceToHost));
    }
    if (fast_pairs) *fast_pairs = (int64_t)h[0];
    if (all_pairs) *all_pairs = (int64_t)h[1];
    return ICQB_OK;
}

}  // extern "C"
