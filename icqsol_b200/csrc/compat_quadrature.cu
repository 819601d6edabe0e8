// The reference's quadrature handle API (bem/icqQuadrature.h:19-40) on top of
// the device arithmetic: same names, argument order and semantics, so that a
// testQuadratureCpp-style known-answer test exercises the GPU code.
//   icqQuadratureEvaluate        bem/icqQuadrature.cpp:176-206
//   icqQuadratureEvaluateDouble  bem/icqQuadrature.cpp:208-246
// Unknown orders return 0 (:192,:227); a zero-area triangle returns 0 from
// Evaluate (:185).  Each call runs one CUDA block and is blocking.
#include <math.h>

#include "icqb_internal.cuh"
#include "quad_tables.cuh"

struct icqQuadratureType {
    double pObs[3];
    double* d_in;    // 21 doubles
    double* d_out;   // 1 double
};

namespace {

using namespace icqb;

__device__ double cross_norm(const double* db, const double* dc) {
    const double p12 = __dmul_rn(db[1], dc[2]), p20 = __dmul_rn(db[2], dc[0]);
    const double p01 = __dmul_rn(db[0], dc[1]), p21 = __dmul_rn(db[2], dc[1]);
    const double p02 = __dmul_rn(db[0], dc[2]), p10 = __dmul_rn(db[1], dc[0]);
    const double cx = __dsub_rn(p12, p21), cy = __dsub_rn(p20, p02), cz = __dsub_rn(p01, p10);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz)));
}

__device__ double node_point(double pa, double db, double dc, double xi, double eta) {
    return __dadd_rn(__dadd_rn(pa, __dmul_rn(xi, db)), __dmul_rn(eta, dc));
}

__device__ double laplace(const double* p, const double* obs) {
    const double dx = p[0] - obs[0], dy = p[1] - obs[1], dz = p[2] - obs[2];
    return 1.0 / sqrt(dx * dx + dy * dy + dz * dz) / (-4.0 * 3.14159265358979323846);
}

// in: obs[3], pa[3], pb[3], pc[3]
__global__ void k_eval_single(int order, const double* in, double* out) {
    if (threadIdx.x != 0) return;
    double res = 0.0;
    const double *obs = in, *pa = in + 3, *pb = in + 6, *pc = in + 9;
    double db[3], dc[3];
    for (int k = 0; k < 3; ++k) {
        db[k] = pb[k] - pa[k];
        dc[k] = pc[k] - pa[k];
    }
    const double area = cross_norm(db, dc);
    if (area == 0 || order < 1 || order > kTriMaxOrder) {
        *out = 0.0;
        return;
    }
    const int o = c_triRuleStart[order - 1], n = c_triRuleStart[order] - o;
    for (int i = 0; i < n; ++i) {
        double p[3];
        for (int k = 0; k < 3; ++k)
            p[k] = node_point(pa[k], db[k], dc[k], c_triNodes[o + i].xi, c_triNodes[o + i].eta);
        res += laplace(p, obs) * c_triNodes[o + i].w;
    }
    *out = 0.5 * area * res;
}

// in: pa0,pb0,pc0,pa1,pb1,pc1 ; one thread per observer node, combined in order
__global__ void k_eval_double(int order, const double* in, double* out) {
    __shared__ double s_row[32];
    const double *pa0 = in, *pb0 = in + 3, *pc0 = in + 6, *pa1 = in + 9, *pb1 = in + 12,
                 *pc1 = in + 15;
    if (order < 1 || order > kTriMaxOrder) {
        if (threadIdx.x == 0) *out = 0.0;
        return;
    }
    const int o = c_triRuleStart[order - 1], n = c_triRuleStart[order] - o;
    double db0[3], dc0[3], db1[3], dc1[3];
    for (int k = 0; k < 3; ++k) {
        db0[k] = pb0[k] - pa0[k];
        dc0[k] = pc0[k] - pa0[k];
        db1[k] = pb1[k] - pa1[k];
        dc1[k] = pc1[k] - pa1[k];
    }
    const int i0 = threadIdx.x;
    double row = 0.0;
    if (i0 < n) {
        double p0[3];
        for (int k = 0; k < 3; ++k)
            p0[k] = node_point(pa0[k], db0[k], dc0[k], c_triNodes[o + i0].xi, c_triNodes[o + i0].eta);
        for (int i1 = 0; i1 < n; ++i1) {
            double p1[3];
            for (int k = 0; k < 3; ++k)
                p1[k] = node_point(pa1[k], db1[k], dc1[k], c_triNodes[o + i1].xi,
                                   c_triNodes[o + i1].eta);
            row += c_triNodes[o + i0].w * laplace(p1, p0) * c_triNodes[o + i1].w;
        }
    }
    s_row[threadIdx.x] = row;
    __syncthreads();
    if (threadIdx.x == 0) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += s_row[i];
        *out = 0.5 * cross_norm(db1, dc1) * res;
    }
}

double run(icqQuadratureType* q, int order, const double* host_in, int count, bool dbl) {
    double result = nan("");
    if (cudaMemcpy(q->d_in, host_in, sizeof(double) * count, cudaMemcpyHostToDevice) != cudaSuccess)
        return result;
    if (dbl)
        k_eval_double<<<1, 32>>>(order, q->d_in, q->d_out);
    else
        k_eval_single<<<1, 32>>>(order, q->d_in, q->d_out);
    if (cudaMemcpy(&result, q->d_out, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        return nan("");
    return result;
}

}  // namespace

extern "C" {

void icqQuadratureInit(icqQuadratureType** self) {
    *self = nullptr;
    icqQuadratureType* q = new icqQuadratureType();
    q->pObs[0] = q->pObs[1] = q->pObs[2] = 0.0;   // icqFunctor.cpp:4-6
    q->d_in = q->d_out = nullptr;
    if (cudaMalloc(&q->d_in, sizeof(double) * 24) != cudaSuccess ||
        cudaMalloc(&q->d_out, sizeof(double)) != cudaSuccess) {
        icqb::set_error(nullptr, ICQB_ECUDA, "icqQuadratureInit: no CUDA device; this library has no CPU path");
        cudaFree(q->d_in);
        delete q;
        return;  // *self stays NULL: callers fail loudly
    }
    *self = q;
}

void icqQuadratureSetObserver(icqQuadratureType** self, const double* pObs) {
    for (int k = 0; k < 3; ++k) (*self)->pObs[k] = pObs[k];
}

int icqQuadratureGetMaxOrder(icqQuadratureType** self) {
    (void)self;
    return icqb::kTriMaxOrder;
}

double icqQuadratureEvaluate(icqQuadratureType** self, int order, const double* pa,
                             const double* pb, const double* pc) {
    double in[12];
    for (int k = 0; k < 3; ++k) {
        in[k] = (*self)->pObs[k];
        in[3 + k] = pa[k];
        in[6 + k] = pb[k];
        in[9 + k] = pc[k];
    }
    return run(*self, order, in, 12, false);
}

double icqQuadratureEvaluateDouble(icqQuadratureType** self, int order, const double* pa0,
                                   const double* pb0, const double* pc0, const double* pa1,
                                   const double* pb1, const double* pc1) {
    double in[18];
    const double* src[6] = {pa0, pb0, pc0, pa1, pb1, pc1};
    for (int v = 0; v < 6; ++v)
        for (int k = 0; k < 3; ++k) in[3 * v + k] = src[v][k];
    return run(*self, order, in, 18, true);
}

void icqQuadratureDel(icqQuadratureType** self) {
    if (!self || !*self) return;
    cudaFree((*self)->d_in);
    cudaFree((*self)->d_out);
    delete *self;
    *self = nullptr;
}

}  // extern "C"
