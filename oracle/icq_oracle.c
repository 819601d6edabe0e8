/* TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Plain-C CPU restatement of icqsol's boundary-element hot path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (icqsol_b200) never does.
 *
 * Parity pin: every function below is checked in tests/test_oracle.py against
 *   - the reference's own C++ (oracle/_ref/libicqref.so, built from the
 *     sources where they lie under /root/reference by oracle/Makefile), and
 *   - golden vectors produced by importing the reference's Python modules
 *     (tests/golden/make_golden.py -> tests/golden/*.npz).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "icq_tables.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* bem/icqQuadrature.cpp:158-161 -- number of tabulated triangle rules. */
int icqo_max_order(void) { return ICQO_TRI_MAX_ORDER; }

/* number of nodes of triangle rule `order`, 0 for an unknown order
 * (bem/icqQuadrature.cpp:192,227: unknown order contributes nothing). */
int icqo_tri_rule(int order, const double** xi, const double** eta, const double** w) {
    if (order < 1 || order > ICQO_TRI_MAX_ORDER) return 0;
    int o = ICQO_TRI_OFF[order - 1];
    *xi = ICQO_TRI_XI + o;
    *eta = ICQO_TRI_ETA + o;
    *w = ICQO_TRI_W + o;
    return ICQO_TRI_OFF[order] - o;
}

/* bem/icqQuadrature.cpp:163-174 and bem/icqLaplaceMatrices.cpp:9-30:
 * "area" is the norm of the cross product, i.e. TWICE the triangle area. */
static double cross_norm(const double* db, const double* dc) {
    const double p12 = db[1] * dc[2];
    const double p20 = db[2] * dc[0];
    const double p01 = db[0] * dc[1];
    const double p21 = db[2] * dc[1];
    const double p02 = db[0] * dc[2];
    const double p10 = db[1] * dc[0];
    return sqrt((p12 - p21) * (p12 - p21) + (p20 - p02) * (p20 - p02) +
                (p01 - p10) * (p01 - p10));
}

double icqo_area2(const double* pa, const double* pb, const double* pc) {
    double db[3], dc[3];
    for (int j = 0; j < 3; ++j) {
        db[j] = pb[j] - pa[j];
        dc[j] = pc[j] - pa[j];
    }
    return cross_norm(db, dc);
}

/* bem/icqLaplaceFunctor.cpp:4-10 */
static double laplace_kernel(const double* pSrc, const double* pObs) {
    double dx = pSrc[0] - pObs[0];
    double dy = pSrc[1] - pObs[1];
    double dz = pSrc[2] - pObs[2];
    return 1.0 / sqrt(dx * dx + dy * dy + dz * dz) / (-4. * M_PI);
}

/* bem/icqQuadrature.cpp:176-206 -- single-triangle quadrature of the Laplace
 * kernel seen from `pobs` (the functor's observer, icqFunctor.cpp:10-15). */
double icqo_quadrature_evaluate(int order, const double* pobs, const double* pa,
                                const double* pb, const double* pc) {
    double res = 0;
    const double pb2[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    const double pc2[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
    double area = cross_norm(pb2, pc2);
    if (area == 0) return res;
    const double *xi, *eta, *w;
    int n = icqo_tri_rule(order, &xi, &eta, &w);
    double p[3];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 3; ++j) p[j] = pa[j] + xi[i] * pb2[j] + eta[i] * pc2[j];
        res += laplace_kernel(p, pobs) * w[i];
    }
    return 0.5 * area * res;
}

/* bem/icqQuadrature.cpp:208-246 -- double quadrature: triangle 0 is the
 * observer (averaged, no area factor), triangle 1 the source. */
double icqo_quadrature_evaluate_double(int order, const double* pa0, const double* pb0,
                                       const double* pc0, const double* pa1,
                                       const double* pb1, const double* pc1) {
    double res = 0;
    const double db0[3] = {pb0[0] - pa0[0], pb0[1] - pa0[1], pb0[2] - pa0[2]};
    const double dc0[3] = {pc0[0] - pa0[0], pc0[1] - pa0[1], pc0[2] - pa0[2]};
    const double db1[3] = {pb1[0] - pa1[0], pb1[1] - pa1[1], pb1[2] - pa1[2]};
    const double dc1[3] = {pc1[0] - pa1[0], pc1[1] - pa1[1], pc1[2] - pa1[2]};
    const double area1 = cross_norm(db1, dc1);
    const double *xi, *eta, *w;
    int n = icqo_tri_rule(order, &xi, &eta, &w);
    double p0[3], p1[3];
    for (int i0 = 0; i0 < n; ++i0) {
        for (int j = 0; j < 3; ++j) p0[j] = pa0[j] + xi[i0] * db0[j] + eta[i0] * dc0[j];
        for (int i1 = 0; i1 < n; ++i1) {
            for (int j = 0; j < 3; ++j) p1[j] = pa1[j] + xi[i1] * db1[j] + eta[i1] * dc1[j];
            res += w[i0] * laplace_kernel(p1, p0) * w[i1];
        }
    }
    return 0.5 * area1 * res;
}

static void tri_vertices(const double* xyz, const long long* tri, long long t,
                         double* pa, double* pb, double* pc) {
    memcpy(pa, xyz + 3 * tri[3 * t + 0], 3 * sizeof(double));
    memcpy(pb, xyz + 3 * tri[3 * t + 1], 3 * sizeof(double));
    memcpy(pc, xyz + 3 * tri[3 * t + 2], 3 * sizeof(double));
}

/* bem/icqLaplaceMatrices.cpp:37-105 -- every pair iObs > jSrc at the maximum
 * order; lower entry direct, upper entry by the area-ratio mirror; diagonal
 * left untouched.  The j loop is the one the reference marks `omp parallel
 * for` (line 74); here the per-thread vertex buffers are private, so enabling
 * OpenMP is race-free (the reference's pragma is inert and would race). */
void icqo_offdiag(const double* xyz, const long long* tri, long long ntri, double* gMat) {
    const int maxOrder = icqo_max_order();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (long long jSrc = 0; jSrc < ntri; ++jSrc) {
        double paS[3], pbS[3], pcS[3], paO[3], pbO[3], pcO[3];
        tri_vertices(xyz, tri, jSrc, paS, pbS, pcS);
        double areaSrc = icqo_area2(paS, pbS, pcS);
        for (long long iObs = jSrc + 1; iObs < ntri; ++iObs) {
            tri_vertices(xyz, tri, iObs, paO, pbO, pcO);
            double areaObs = icqo_area2(paO, pbO, pcO);
            double g = icqo_quadrature_evaluate_double(maxOrder, paO, pbO, pcO, paS, pbS, pcS);
            gMat[ntri * iObs + jSrc] = g;
            gMat[ntri * jSrc + iObs] = g * areaObs / areaSrc;
        }
    }
}

/* One entry G[r,c], r != c, exactly as icqo_offdiag would store it
 * (bem/icqLaplaceMatrices.cpp:94-98): lets tests sample a huge matrix. */
double icqo_offdiag_entry(const double* xyz, const long long* tri, long long r, long long c) {
    double paS[3], pbS[3], pcS[3], paO[3], pbO[3], pcO[3];
    long long iObs = r > c ? r : c, jSrc = r > c ? c : r;
    tri_vertices(xyz, tri, jSrc, paS, pbS, pcS);
    tri_vertices(xyz, tri, iObs, paO, pbO, pcO);
    double g = icqo_quadrature_evaluate_double(icqo_max_order(), paO, pbO, pcO, paS, pbS, pcS);
    if (r > c) return g;
    return g * icqo_area2(paO, pbO, pcO) / icqo_area2(paS, pbS, pcS);
}

/* Rectangular block out[(r-r0)*(c1-c0) + (c-c0)] = G[r,c] for r in [r0,r1),
 * c in [c0,c1); diagonal positions are left untouched. */
void icqo_offdiag_block(const double* xyz, const long long* tri, long long r0, long long r1,
                        long long c0, long long c1, double* out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (long long r = r0; r < r1; ++r)
        for (long long c = c0; c < c1; ++c)
            if (r != c) out[(r - r0) * (c1 - c0) + (c - c0)] = icqo_offdiag_entry(xyz, tri, r, c);
}

/* bem/icqReferenceTriangle.py:14-45 + bem/icqPotentialIntegrals.py:15-40 +
 * bem/icqQuadrature1D.py:28-40 -- integral of 1/R over the triangle
 * (x, p, q) with the observer at corner x: analytic in r, `order`-point
 * Gauss-Legendre in the polar angle.  Returns NaN for an unknown order (the
 * reference raises KeyError, icqQuadrature1D.py:38). */
double icqo_corner_integral(int order, const double* x, const double* p, const double* q) {
    if (order < 1 || order > ICQO_LINE_MAX_ORDER) return NAN;
    double b[3], c[3], n[3];
    for (int j = 0; j < 3; ++j) {
        b[j] = p[j] - x[j];
        c[j] = q[j] - x[j];
    }
    n[0] = b[1] * c[2] - b[2] * c[1];
    n[1] = b[2] * c[0] - b[0] * c[2];
    n[2] = b[0] * c[1] - b[1] * c[0];
    const double r1 = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    const double r2 = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const double ncross = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double bigTheta = atan2(ncross, b[0] * c[0] + b[1] * c[1] + b[2] * c[2]);
    const double a = (r2 * cos(bigTheta) - r1) / r2 / sin(bigTheta);
    const int o = ICQO_LINE_OFF[order - 1], m = ICQO_LINE_OFF[order] - o;
    double s = 0;
    for (int k = 0; k < m; ++k) {
        double t = 0.0 + ICQO_LINE_U[o + k] * bigTheta;
        double v = ICQO_LINE_W[o + k] * (r1 / (cos(t) - a * sin(t)));
        s = (k == 0) ? v : s + v;
    }
    return bigTheta * s;
}

/* bem/icqBaseLaplaceSolver.py:34-66 -- self terms G[j,j]; `order` selects
 * both the triangle rule for the observer points and the line rule. */
int icqo_diag(const double* xyz, const long long* tri, long long ntri, int order, double* diag) {
    if (order < 1 || order > ICQO_LINE_MAX_ORDER) return 1;
    const double *xi, *eta, *w;
    int n = icqo_tri_rule(order, &xi, &eta, &w);
    const double fourPi = 4. * M_PI;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (long long t = 0; t < ntri; ++t) {
        double pa[3], pb[3], pc[3], db[3], dc[3], xo[3];
        tri_vertices(xyz, tri, t, pa, pb, pc);
        for (int j = 0; j < 3; ++j) {
            db[j] = pb[j] - pa[j];
            dc[j] = pc[j] - pa[j];
        }
        double g = 0;
        for (int k = 0; k < n; ++k) {
            for (int j = 0; j < 3; ++j) xo[j] = pa[j] + xi[k] * db[j] + eta[k] * dc[j];
            g += w[k] * (icqo_corner_integral(order, xo, pa, pb) +
                         icqo_corner_integral(order, xo, pb, pc) +
                         icqo_corner_integral(order, xo, pc, pa));
        }
        diag[t] = g / (-fourPi);
    }
    return 0;
}

/* K (double layer) -- NOT IN THE REFERENCE SNAPSHOT:
 * PARITY UNPINNED.  Definition used by the product, with G's conventions
 * (bem/icqQuadrature.cpp:208-246): observer-averaged, source-integrated
 *   K[i,j] = 0.5*A2_j * sum_i0 sum_i1 w_i0 w_i1 * (-1/(4 pi)) * n_j.(x-y)/|x-y|^3
 * with x on observer triangle i, y on source triangle j and
 * n_j = (db x dc)/|db x dc| from the vertex order; K[i,i] = 0. */
double icqo_k_entry(const double* xyz, const long long* tri, long long i, long long j) {
    if (i == j) return 0.0;
    double pa0[3], pb0[3], pc0[3], pa1[3], pb1[3], pc1[3];
    tri_vertices(xyz, tri, i, pa0, pb0, pc0);
    tri_vertices(xyz, tri, j, pa1, pb1, pc1);
    double db0[3], dc0[3], db1[3], dc1[3], nrm[3];
    for (int k = 0; k < 3; ++k) {
        db0[k] = pb0[k] - pa0[k];
        dc0[k] = pc0[k] - pa0[k];
        db1[k] = pb1[k] - pa1[k];
        dc1[k] = pc1[k] - pa1[k];
    }
    nrm[0] = db1[1] * dc1[2] - db1[2] * dc1[1];
    nrm[1] = db1[2] * dc1[0] - db1[0] * dc1[2];
    nrm[2] = db1[0] * dc1[1] - db1[1] * dc1[0];
    const double area1 = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    const double *xi, *eta, *w;
    int n = icqo_tri_rule(icqo_max_order(), &xi, &eta, &w);
    double res = 0;
    for (int i0 = 0; i0 < n; ++i0) {
        double x[3];
        for (int k = 0; k < 3; ++k) x[k] = pa0[k] + xi[i0] * db0[k] + eta[i0] * dc0[k];
        for (int i1 = 0; i1 < n; ++i1) {
            double d[3];
            for (int k = 0; k < 3; ++k)
                d[k] = x[k] - (pa1[k] + xi[i1] * db1[k] + eta[i1] * dc1[k]);
            double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double r = sqrt(r2);
            double nd = (nrm[0] * d[0] + nrm[1] * d[1] + nrm[2] * d[2]) / area1;
            res += w[i0] * w[i1] * nd / (r2 * r) / (-4. * M_PI);
        }
    }
    return 0.5 * area1 * res;
}

void icqo_k_block(const double* xyz, const long long* tri, long long r0, long long r1,
                  long long c0, long long c1, double* out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (long long r = r0; r < r1; ++r)
        for (long long c = c0; c < c1; ++c)
            out[(r - r0) * (c1 - c0) + (c - c0)] = icqo_k_entry(xyz, tri, r, c);
}

/* y = A x, row-major dense (numpy.dot at bem/icqPoissonSolver.py:36). */
void icqo_gemv(const double* a, long long nrows, long long ncols, const double* x, double* y) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (long long i = 0; i < nrows; ++i) {
        double s = 0;
        const double* row = a + i * ncols;
        for (long long j = 0; j < ncols; ++j) s += row[j] * x[j];
        y[i] = s;
    }
}

int icqo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
