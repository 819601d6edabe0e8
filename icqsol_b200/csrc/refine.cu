// Uniform refinement on the device: every triangle is cut into k*k children (k per edge),
// children keep the parent's orientation.
//
// This is the part of the reference's RefineSurface (shapes/icqRefineSurface.py:46-177:
// edge splitting, then a Delaunay triangulation by pytriangle, cell data copied from the
// parent to its children, :155-174) that is reproducible without pytriangle: it produces
// the refined meshes of BASELINE.json's large configurations (e.g. test_results/
// testBoltFine.vtk, 2,350 triangles, k = 6 -> 84,600) inside the run.  Same point and
// child order and the same roundings as the host version
// icqsol_b200.mesh.generators.subdivide, which is its bit-exact check:
//   points of parent t:  lattice (i, j), i + j <= k, ordered by i then j,
//                        a + (i/k)(b - a) + (j/k)(c - a), optionally rounded to float32
//                        (VTK's default point type, shapes/icqShapeManager.py:794-798);
//   children of parent t: for i, for j: (i,j),(i+1,j),(i,j+1) and, if i + j < k - 1,
//                        (i+1,j),(i+1,j+1),(i,j+1).
// Points on shared edges are duplicated per parent (G does not need a conforming mesh).
#include <algorithm>

#include "icqb_internal.cuh"

namespace icqb {

namespace {

// first lattice index of row i: sum_{i' < i} (k + 1 - i')
__device__ __forceinline__ long long lattice_row(long long i, long long k) {
    return i * (k + 1) - i * (i - 1) / 2;
}

__global__ void __launch_bounds__(256) k_subdivide(const double* __restrict__ xyz,
                                                   const int64_t* __restrict__ tri, long long ntri,
                                                   int k, int f32, double* __restrict__ xyzOut,
                                                   int64_t* __restrict__ triOut) {
    const long long npl = (long long)(k + 1) * (k + 2) / 2;   // points per parent
    const long long nch = (long long)k * k;                   // children per parent
    const long long per = npl + nch;
    const long long total = ntri * per;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / per, q = e - t * per;
        if (q < npl) {
            // lattice point q -> (i, j)
            long long i = 0;
            while (lattice_row(i + 1, k) <= q) ++i;
            const long long j = q - lattice_row(i, k);
            const double iu = (double)i / (double)k, ju = (double)j / (double)k;
            const long long ia = tri[3 * t + 0], ib = tri[3 * t + 1], ic = tri[3 * t + 2];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double a = xyz[3 * ia + d], b = xyz[3 * ib + d], c = xyz[3 * ic + d];
                double p = __dadd_rn(__dadd_rn(a, __dmul_rn(iu, __dsub_rn(b, a))),
                                     __dmul_rn(ju, __dsub_rn(c, a)));
                if (f32) p = (double)(float)p;
                xyzOut[3 * (t * npl + q) + d] = p;
            }
        } else {
            // child c -> row i, position in the row, upright / inverted
            const long long c = q - npl;
            long long i = 0;
            while (2 * k * (i + 1) - (i + 1) * (i + 1) <= c) ++i;   // row i starts at 2ki - i^2
            const long long pos = c - (2 * k * i - i * i);
            const long long j = pos / 2;
            const long long base = t * npl;
            const long long r0 = lattice_row(i, k), r1 = lattice_row(i + 1, k);
            long long v0, v1, v2;
            if ((pos & 1) == 0) {
                v0 = r0 + j; v1 = r1 + j; v2 = r0 + j + 1;
            } else {
                v0 = r1 + j; v1 = r1 + j + 1; v2 = r0 + j + 1;
            }
            triOut[3 * (t * nch + c) + 0] = base + v0;
            triOut[3 * (t * nch + c) + 1] = base + v1;
            triOut[3 * (t * nch + c) + 2] = base + v2;
        }
    }
}

}  // namespace

}  // namespace icqb

using namespace icqb;

extern "C" {

int icqb_subdivide_dev(icqb_ctx* ctx, uint64_t dxyz, uint64_t dtri, int64_t ntri, int k,
                       int float32_points, uint64_t dxyz_out, uint64_t dtri_out) {
    if (!ctx) return ICQB_EARG;
    if (k < 1 || ntri < 0) return set_error(ctx, ICQB_EARG, "icqb_subdivide_dev: need k >= 1, ntri >= 0");
    if (ntri == 0) return ICQB_OK;
    if (!dxyz || !dtri || !dxyz_out || !dtri_out)
        return set_error(ctx, ICQB_EARG, "icqb_subdivide_dev: null pointer");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long per = (long long)(k + 1) * (k + 2) / 2 + (long long)k * k;
    const long long total = (long long)ntri * per;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
    k_subdivide<<<blocks, 256, 0, ctx->stream>>>(
        reinterpret_cast<const double*>(dxyz), reinterpret_cast<const int64_t*>(dtri), ntri, k,
        float32_points, reinterpret_cast<double*>(dxyz_out), reinterpret_cast<int64_t*>(dtri_out));
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_subdivide launch");
}

int icqb_subdivide_host(icqb_ctx* ctx, const double* xyz, int64_t npts, const int64_t* tri,
                        int64_t ntri, int k, int float32_points, double* xyz_out, int64_t* tri_out) {
    if (!ctx) return ICQB_EARG;
    if (k < 1 || ntri < 0 || npts < 0)
        return set_error(ctx, ICQB_EARG, "icqb_subdivide_host: need k >= 1 and non-negative sizes");
    if (ntri == 0) return ICQB_OK;
    if (!xyz || !tri || !xyz_out || !tri_out)
        return set_error(ctx, ICQB_EARG, "icqb_subdivide_host: null pointer");
    for (int64_t e = 0; e < 3 * ntri; ++e)
        if (tri[e] < 0 || tri[e] >= npts)
            return set_error(ctx, ICQB_EARG, "icqb_subdivide_host: vertex index out of range");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long npl = (long long)(k + 1) * (k + 2) / 2, nch = (long long)k * k;
    double *dxyz = nullptr, *dout = nullptr;
    int64_t *dtri = nullptr, *dtout = nullptr;
    int st = ICQB_OK;
    auto cleanup = [&]() {
        cudaFree(dxyz);
        cudaFree(dout);
        cudaFree(dtri);
        cudaFree(dtout);
    };
#define SUB_TRY(call)                                         \
    do {                                                      \
        st = check_cuda(ctx, (call), #call);                  \
        if (st != ICQB_OK) {                                  \
            cleanup();                                        \
            return st;                                        \
        }                                                     \
    } while (0)
    SUB_TRY(cudaMalloc(&dxyz, sizeof(double) * 3 * (npts > 0 ? npts : 1)));
    SUB_TRY(cudaMalloc(&dtri, sizeof(int64_t) * 3 * ntri));
    SUB_TRY(cudaMalloc(&dout, sizeof(double) * 3 * ntri * npl));
    SUB_TRY(cudaMalloc(&dtout, sizeof(int64_t) * 3 * ntri * nch));
    SUB_TRY(cudaMemcpyAsync(dxyz, xyz, sizeof(double) * 3 * npts, cudaMemcpyHostToDevice, ctx->stream));
    SUB_TRY(cudaMemcpyAsync(dtri, tri, sizeof(int64_t) * 3 * ntri, cudaMemcpyHostToDevice, ctx->stream));
    st = icqb_subdivide_dev(ctx, (uint64_t)dxyz, (uint64_t)dtri, ntri, k, float32_points,
                            (uint64_t)dout, (uint64_t)dtout);
    if (st != ICQB_OK) {
        cleanup();
        return st;
    }
    SUB_TRY(cudaMemcpyAsync(xyz_out, dout, sizeof(double) * 3 * ntri * npl, cudaMemcpyDeviceToHost,
                            ctx->stream));
    SUB_TRY(cudaMemcpyAsync(tri_out, dtout, sizeof(int64_t) * 3 * ntri * nch, cudaMemcpyDeviceToHost,
                            ctx->stream));
    SUB_TRY(cudaStreamSynchronize(ctx->stream));
#undef SUB_TRY
    cleanup();
    return ICQB_OK;
}

}  // extern "C"
