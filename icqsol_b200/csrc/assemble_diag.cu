// Self terms G[j,j]: integral of 1/R over the triangle seen from each of its own
// quadrature points, split into three corner triangles and done in polar form --
// analytic in r, `order`-point Gauss-Legendre in the angle.
//
// Replaces bem/icqBaseLaplaceSolver.py:34-66 with bem/icqPotentialIntegrals.py:15-40,
// bem/icqReferenceTriangle.py:14-45 and bem/icqQuadrature1D.py:28-40.  NOTE: this
// reproduces the reference's *quadrature* of the angular integral, not the exact
// closed form -- at order 5 they differ by ~4e-4 relative.
//
// Mapping: 8 lanes per triangle (the order-5 rule has 7 observer points; lanes
// beyond the rule's node count idle), 4 triangles per warp; the per-point terms
// are combined in the reference's sequential order by the group's first lane.
#include <math.h>

#include "icqb_internal.cuh"
#include "quad_tables.cuh"

namespace icqb {

namespace {

// PotentialIntegrals(x, p, q, order).getIntegralOneOverR()
__device__ double corner_integral(int order, const double* x, const double* p, const double* q) {
    double b[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b[k] = p[k] - x[k];
        c[k] = q[k] - x[k];
    }
    const double nx = b[1] * c[2] - b[2] * c[1];
    const double ny = b[2] * c[0] - b[0] * c[2];
    const double nz = b[0] * c[1] - b[1] * c[0];
    const double r1 = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    const double r2 = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const double ncross = sqrt(nx * nx + ny * ny + nz * nz);
    const double bigTheta = atan2(ncross, b[0] * c[0] + b[1] * c[1] + b[2] * c[2]);
    double sT, cT;
    sincos(bigTheta, &sT, &cT);
    const double a = (r2 * cT - r1) / r2 / sT;
    const int o = c_lineRuleStart[order - 1], m = c_lineRuleStart[order] - o;
    double s = 0.0;
    for (int k = 0; k < m; ++k) {
        const double t = c_lineNodes[o + k].u * bigTheta;
        double st, ct;
        sincos(t, &st, &ct);
        s += c_lineNodes[o + k].w * (r1 / (ct - a * st));
    }
    return bigTheta * s;
}

__global__ void __launch_bounds__(128) k_diag(const double* __restrict__ verts, long long r0,
                                              long long r1, int order, double* __restrict__ out,
                                              long long stride) {
    const int lane8 = threadIdx.x & 7;
    const long long t = r0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 8;
    const bool valid = t < r1;
    const long long tl = valid ? t : (r1 - 1);
    const int o = c_triRuleStart[order - 1], npts = c_triRuleStart[order] - o;

    double pa[3], pb[3], pc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pa[k] = __ldg(verts + 9 * tl + k);
        pb[k] = __ldg(verts + 9 * tl + 3 + k);
        pc[k] = __ldg(verts + 9 * tl + 6 + k);
    }
    double term = 0.0;
    if (lane8 < npts) {
        const double xi = c_triNodes[o + lane8].xi, eta = c_triNodes[o + lane8].eta;
        double x[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // xObs = paSrc + xsis*dbSrc + etas*dcSrc, unfused (icqBaseLaplaceSolver.py:57)
            const double db = __dsub_rn(pb[k], pa[k]), dc = __dsub_rn(pc[k], pa[k]);
            x[k] = __dadd_rn(__dadd_rn(pa[k], __dmul_rn(xi, db)), __dmul_rn(eta, dc));
        }
        const double iab = corner_integral(order, x, pa, pb);
        const double ibc = corner_integral(order, x, pb, pc);
        const double ica = corner_integral(order, x, pc, pa);
        term = c_triNodes[o + lane8].w * ((iab + ibc) + ica);
    }
    // g = sum over observer points in table order (icqBaseLaplaceSolver.py:62-64)
    double g = 0.0;
    const unsigned full = 0xffffffffu;
    const int base = (threadIdx.x & 31) & ~7;
    for (int k = 0; k < 7; ++k) {
        const double v = __shfl_sync(full, term, base + k);
        if (k < npts) g += v;
    }
    if (valid && lane8 == 0) out[(t - r0) * stride] = g / (-4.0 * 3.14159265358979323846);
}

}  // namespace

// self terms of global rows [g0, g1): out[(t - g0)*stride]
int launch_diag_rows(icqb_ctx* ctx, int order, long long g0, long long g1, double* dOut,
                     int64_t stride) {
    if (order < 1 || order > kLineMaxOrder)
        return set_error(ctx, ICQB_EORDER, "diagonal order must be in 1..5 (KeyError in the reference)");
    const long long rows = g1 - g0;
    if (rows <= 0) return ICQB_OK;
    const int threads = 128;
    const unsigned blocks = (unsigned)((rows * 8 + threads - 1) / threads);
    k_diag<<<blocks, threads, 0, ctx->stream>>>(ctx->d_verts, g0, g1, order, dOut, stride);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_diag launch");
}

int launch_diag(icqb_ctx* ctx, int order, double* dOut, int64_t stride) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    const int st = launch_diag_rows(ctx, order, ctx->r0, ctx->r1, dOut, stride);
    cudaEventRecord(ctx->ev1, ctx->stream);
    return st;
}

}  // namespace icqb
