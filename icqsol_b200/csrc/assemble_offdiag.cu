// Fused off-diagonal assembly of the single-layer (G) and double-layer (K)
// Laplace matrices over triangle pairs, fp64, sm_100a.
//
// Replaces bem/icqLaplaceMatrices.cpp:75-101 (pair loop, area-ratio mirror),
// bem/icqQuadrature.cpp:208-246 (16x16-point double quadrature) and
// bem/icqLaplaceFunctor.cpp:4-10 (1/(-4 pi r)) by one kernel.
//
// Mapping (DESIGN.md "assemble_offdiag"):
//   * one CTA = 128 observer triangles (one per thread) x 128 source triangles;
//   * the thread keeps its observer triangle's 16 quadrature points in registers
//     (48 doubles), weights are constant-bank operands of the DFMAs;
//   * source triangles arrive as 32-triangle sub-tiles (32 x 384 B, contiguous in
//     the triangle-major table) through the bulk-copy engine
//     (cp.async.bulk ... mbarrier::complete_tx) into a 2-stage shared-memory ring;
//     every lane reads the same source point -> shared-memory broadcast;
//   * per evaluation: 3 DADD + DMUL + 2 DFMA (r^2), MUFU.RSQ64H + 5 FP64 (1/r to
//     <= 1 ulp, branch free), 1 DFMA (weighted accumulate) = 12 FP64-pipe
//     instructions; K adds 2 DMUL + 2 DFMA (1/r^3 into a row and a column sum);
//   * the pair sum S_ij = sum w_i0 w_i1 / r is symmetric, so inside the locally
//     owned square block only tiles with I >= J are evaluated and both
//     G[i,j] = c A_j S and G[j,i] = c A_i S are stored (the reference's
//     g*areaObs/areaSrc mirror, icqLaplaceMatrices.cpp:97-98); column ranges
//     owned by other ranks are evaluated directly.
//   * the mirrored store is coalesced (lanes = consecutive columns); the direct
//     store is one 8-byte column per lane -- at ~1.5 kflop per entry the store
//     path is idle and L2 merges the sectors, so no staging pass is spent on it.
#include "offdiag_common.cuh"

namespace icqb {

namespace {

template <bool WITH_K>
__global__ void __launch_bounds__(kTileObs) k_offdiag(const OffDiagParams P) {
    __shared__ __align__(128) double s_src[kStages][kSub * kQpStride];
    __shared__ __align__(8) uint64_t s_bar[kStages];

    const int tid = threadIdx.x;
    const Tile T = decode_tile(P, blockIdx.x);
    if (T.kind < 0) return;  // lower storage: tile above the diagonal (whole CTA exits)

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
#ifndef ICQB_EMU
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();

    // sub-tiles of this CTA that hold at least one valid source triangle; only
    // those are ever copied, and every issued copy is waited on before exit
    const long long srcLeft = T.srcEnd - T.src0;
    const int nSubTiles =
        (int)(srcLeft >= kTileSrc ? kTileSrc / kSub : (srcLeft + kSub - 1) / kSub);
    if (tid == 0) {
        // prologue: fill the stages (the table is padded by one tile, so a
        // sub-tile that runs past the last triangle still reads valid memory)
        for (int s = 0; s < kStages && s < nSubTiles; ++s) {
            mbar_expect_tx(&s_bar[s], kSubBytes);
            bulk_g2s(&s_src[s][0], P.qp_aos + (T.src0 + (long long)s * kSub) * kQpStride, kSubBytes,
                     &s_bar[s]);
        }
    }

    // observer triangle of this thread
    const long long iObs = T.obs0 + tid;
    const bool obsValid = iObs < T.obsEnd;
    const long long iLoad = obsValid ? iObs : (T.obsEnd - 1);
    double ox[kNq], oy[kNq], oz[kNq];
#pragma unroll
    for (int k = 0; k < kNq; ++k) {
        ox[k] = __ldg(P.qp_soa + (long long)(3 * k + 0) * P.npad + iLoad);
        oy[k] = __ldg(P.qp_soa + (long long)(3 * k + 1) * P.npad + iLoad);
        oz[k] = __ldg(P.qp_soa + (long long)(3 * k + 2) * P.npad + iLoad);
    }
    const double a2Obs = __ldg(P.area2 + iLoad);
    // -1/(4 pi) * 0.5  (icqLaplaceFunctor.cpp:9, icqQuadrature.cpp:245)
    const double cG = 0.5 / (-4.0 * 3.14159265358979323846);

    double nix = 0, niy = 0, niz = 0, pax = 0, pay = 0, paz = 0;
    if (WITH_K) {
        nix = __ldg(P.nrm + 4 * iLoad + 0);
        niy = __ldg(P.nrm + 4 * iLoad + 1);
        niz = __ldg(P.nrm + 4 * iLoad + 2);
        pax = __ldg(P.verts + 9 * iLoad + 0);
        pay = __ldg(P.verts + 9 * iLoad + 1);
        paz = __ldg(P.verts + 9 * iLoad + 2);
    }

    // storage row of this observer, indexed by the GLOBAL column j (clamped lanes never store):
    // full mode  row-major rows of pitch ld;  lower mode  row `tid` of the contiguous tile
    const long long rowOff = P.lower
        ? T.tile * kLowerTileElems + (long long)(obsValid ? tid : 0) * kLowerTile - T.src0
        : (T.lrow0 + (obsValid ? tid : 0)) * P.ld;
    double* gRow = P.G + rowOff;
    double* kRow = WITH_K ? (P.K + rowOff) : nullptr;
    double* kuRow = (WITH_K && P.lower) ? (P.Ku + rowOff) : nullptr;

    for (int sub = 0; sub < nSubTiles; ++sub) {
        const int stage = sub & 1;
        const uint32_t parity = (sub >> 1) & 1;
        const long long j0 = T.src0 + (long long)sub * kSub;
        mbar_wait(&s_bar[stage], parity);
        const double* sp = &s_src[stage][0];

        for (int jj = 0; jj < kSub; ++jj) {
            const long long j = j0 + jj;
            if (j >= T.srcEnd) break;  // uniform
            const double* pj = sp + jj * kQpStride;
            double accG = 0.0;
            double r3row[WITH_K ? kNq : 1];
            double accKm = 0.0;  // mirrored K[j,i]
            double njx = 0, njy = 0, njz = 0, pjax = 0, pjay = 0, pjaz = 0;
            if (WITH_K) {
#pragma unroll
                for (int k = 0; k < kNq; ++k) r3row[k] = 0.0;
                njx = __ldg(P.nrm + 4 * j + 0);
                njy = __ldg(P.nrm + 4 * j + 1);
                njz = __ldg(P.nrm + 4 * j + 2);
                pjax = __ldg(P.verts + 9 * j + 0);
                pjay = __ldg(P.verts + 9 * j + 1);
                pjaz = __ldg(P.verts + 9 * j + 2);
            }
#pragma unroll 1
            for (int i1 = 0; i1 < kNq; ++i1) {
                const double sx = pj[3 * i1 + 0];
                const double sy = pj[3 * i1 + 1];
                const double sz = pj[3 * i1 + 2];
                const double w1 = c_triNodes[kTri8 + i1].w;
                double in0 = 0.0, in1 = 0.0, in2 = 0.0, in3 = 0.0;
                double c30 = 0.0, c31 = 0.0;
#pragma unroll
                for (int i0 = 0; i0 < kNq; ++i0) {
                    const double dx = sx - ox[i0];
                    const double dy = sy - oy[i0];
                    const double dz = sz - oz[i0];
                    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    const double ri = rsqrt_1ulp(r2);
                    const double w0 = c_triNodes[kTri8 + i0].w;
                    if ((i0 & 3) == 0) in0 = fma(w0, ri, in0);
                    if ((i0 & 3) == 1) in1 = fma(w0, ri, in1);
                    if ((i0 & 3) == 2) in2 = fma(w0, ri, in2);
                    if ((i0 & 3) == 3) in3 = fma(w0, ri, in3);
                    if (WITH_K) {
                        const double r3 = ri * ri * ri;
                        r3row[i0] = fma(w1, r3, r3row[i0]);
                        if (i0 & 1) c31 = fma(w0, r3, c31);
                        else c30 = fma(w0, r3, c30);
                    }
                }
                accG = fma(w1, (in0 + in1) + (in2 + in3), accG);
                if (WITH_K) {
                    // mirrored entry K[j,i]: observer point on j (= this source point),
                    // source triangle i: n_i . (x - y) = n_i . (p1 - pa_i)
                    const double nd = fma(niz, sz - paz, fma(niy, sy - pay, nix * (sx - pax)));
                    accKm = fma(w1 * nd, c30 + c31, accKm);
                }
            }

            const bool offDiag = (T.kind != 0) || (j != iObs);
            if (obsValid && offDiag) {
                const double a2Src = __ldg(P.area2 + j);
                gRow[j] = (cG * a2Src) * accG;
                if (T.kind == 1) {
                    // mirror: G[j,i] = g * areaObs / areaSrc  (icqLaplaceMatrices.cpp:98)
                    P.G[(j - P.r0) * P.ld + iObs] = (cG * a2Obs) * accG;
                }
                if (WITH_K) {
                    // K[i,j]: n_j . (x - y) = n_j . (p0 - pa_j)
                    double accK = 0.0;
#pragma unroll
                    for (int i0 = 0; i0 < kNq; ++i0) {
                        const double nd =
                            fma(njz, oz[i0] - pjaz, fma(njy, oy[i0] - pjay, njx * (ox[i0] - pjax)));
                        accK = fma(c_triNodes[kTri8 + i0].w * nd, r3row[i0], accK);
                    }
                    kRow[j] = (cG * a2Src) * accK;
                    if (T.kind == 1) P.K[(j - P.r0) * P.ld + iObs] = (cG * a2Obs) * accKm;
                    if (P.lower) kuRow[j] = (cG * a2Obs) * accKm;   // K[j,i] kept next to K[i,j]
                }
            } else if (obsValid && WITH_K) {
                kRow[j] = 0.0;  // K[i,i] = 0 (flat panel, principal value)
                if (P.lower) kuRow[j] = 0.0;
            }
        }

        // refill this stage with sub-tile sub+2 once every thread is done reading it
        if (sub + kStages < nSubTiles) {
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&s_bar[stage], kSubBytes);
                bulk_g2s(&s_src[stage][0],
                         P.qp_aos + (T.src0 + (long long)(sub + kStages) * kSub) * kQpStride,
                         kSubBytes, &s_bar[stage]);
            }
        }
    }
}

}  // namespace

int launch_offdiag_lower(icqb_ctx* ctx, double* dG, double* dK, double* dKu) {
    if (ctx->ntri == 0) return ICQB_OK;
    OffDiagParams P;
    P.qp_soa = ctx->d_qp_soa;
    P.qp_aos = ctx->d_qp_aos;
    P.area2 = ctx->d_area2;
    P.nrm = ctx->d_nrm;
    P.verts = ctx->d_verts;
    P.G = dG;
    P.K = dK;
    P.Ku = dKu;
    P.ld = 0;
    P.n = ctx->ntri;
    P.npad = ctx->ntri_pad;
    P.r0 = P.r1 = 0;
    P.nObsTiles = P.nLeft = P.nRight = 0;
    P.lower = 1;
    int st = lower_layout(ctx, &P.L);
    if (st != ICQB_OK) return st;
    const long long nLocal = P.L.nLocal();
    if (nLocal == 0) return ICQB_OK;
    // one CTA per stored tile (the 2-D grid of round 1 launched as many CTAs again above the
    // diagonal, only to exit)
    const long long nStored = P.L.numTiles();
    if (nStored > 0x7fffffffLL) return set_error(ctx, ICQB_EARG, "too many tiles for one launch");
    dim3 grid((unsigned)nStored);
    if (ctx->offdiag_variant != 0) return launch_offdiag_ff(ctx, P, grid, dK != nullptr);
    cudaEventRecord(ctx->ev0, ctx->stream);
    if (dK)
        k_offdiag<true><<<grid, kTileObs, 0, ctx->stream>>>(P);
    else
        k_offdiag<false><<<grid, kTileObs, 0, ctx->stream>>>(P);
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_offdiag(lower) launch");
}

int launch_offdiag(icqb_ctx* ctx, double* dG, double* dK, int64_t ld) {
    if (ctx->ntri == 0 || ctx->r1 <= ctx->r0) return ICQB_OK;
    OffDiagParams P;
    P.Ku = nullptr;
    P.L = LowerLayout{0, 0, 0, 0, 0, 0, 0};
    P.lower = 0;
    P.qp_soa = ctx->d_qp_soa;
    P.qp_aos = ctx->d_qp_aos;
    P.area2 = ctx->d_area2;
    P.nrm = ctx->d_nrm;
    P.verts = ctx->d_verts;
    P.G = dG;
    P.K = dK;
    P.ld = ld;
    P.n = ctx->ntri;
    P.npad = ctx->ntri_pad;
    P.r0 = ctx->r0;
    P.r1 = ctx->r1;
    const long long rows = ctx->r1 - ctx->r0;
    P.nObsTiles = (rows + kTileObs - 1) / kTileObs;
    P.nLeft = (ctx->r0 + kTileSrc - 1) / kTileSrc;
    P.nRight = (ctx->ntri - ctx->r1 + kTileSrc - 1) / kTileSrc;
    const long long tasks =
        P.nObsTiles * (P.nObsTiles + 1) / 2 + P.nObsTiles * (P.nLeft + P.nRight);
    if (tasks > 0x7fffffffLL) return set_error(ctx, ICQB_EARG, "too many tiles for one launch");
    if (ctx->offdiag_variant != 0)
        return launch_offdiag_ff(ctx, P, dim3((unsigned)tasks), dK != nullptr);
    cudaEventRecord(ctx->ev0, ctx->stream);
    if (dK)
        k_offdiag<true><<<(unsigned)tasks, kTileObs, 0, ctx->stream>>>(P);
    else
        k_offdiag<false><<<(unsigned)tasks, kTileObs, 0, ctx->stream>>>(P);
    cudaEventRecord(ctx->ev1, ctx->stream);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_offdiag launch");
}

}  // namespace icqb
