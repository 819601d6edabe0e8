// C-ABI entry points that tie the kernels together: assembly (device and host
// buffer forms), the computeOffDiagonalTerms drop-in, GEMV wrappers and the FP64
// peak probe.  See include/icqb.h for the contract of each function.
#include <vector>

#include "icqb_internal.cuh"

using namespace icqb;

namespace {

// register-resident DFMA chains: the denominator of the assembly roofline
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void k_reciprocal(const double* a, double* out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = 1.0 / a[i];
}

// compact self terms (row order of the two blocks) -> diagonal entries of the diagonal tiles
__global__ void k_scatter_diag_lower(const double* diag, LowerLayout L, double* G) {
    const long long rows = (L.r1 - L.r0) + (L.q1 - L.q0);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < rows;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k < (L.r1 - L.r0) ? L.r0 + k : L.q0 + (k - (L.r1 - L.r0));
        const int lt = L.stripOf(i);
        const long long I = i / kLowerTile, r = i - I * kLowerTile;
        G[(L.tileBase(lt) + I) * kLowerTileElems + r * kLowerTile + r] = diag[k];
    }
}

int launch_reciprocal(icqb_ctx* ctx, const double* a, double* out, long long n) {
    k_reciprocal<<<128, 256, 0, ctx->stream>>>(a, out, n);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_reciprocal launch");
}

int finish_phase(icqb_ctx* ctx, int phase) {
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms[phase] = ms;
    return ICQB_OK;
}

}  // namespace

extern "C" {

int icqb_assemble_dev(icqb_ctx* ctx, int diag_order, uint64_t dG, uint64_t dK, int64_t ld) {
    if (!ctx) return ICQB_EARG;
    if (ctx->ntri == 0) return ICQB_OK;
    if (!ctx->d_qp_aos) return set_error(ctx, ICQB_ESTATE, "icqb_assemble_dev: no mesh set");
    if (!dG || ld < ctx->ntri)
        return set_error(ctx, ICQB_EARG, "icqb_assemble_dev: null G or ld < ntri");
    if (diag_order != 0 && (diag_order < 1 || diag_order > 5))
        return set_error(ctx, ICQB_EORDER,
                         "diagonal order must be in 1..5 (KeyError in the reference)");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    double* G = reinterpret_cast<double*>(dG);
    double* K = reinterpret_cast<double*>(dK);
    // reference order: diagonal first, then off-diagonal (icqBaseLaplaceSolver.py:76-77)
    if (diag_order != 0) {
        int st = launch_diag(ctx, diag_order, G + ctx->r0, ld + 1);
        if (st != ICQB_OK) return st;
        st = finish_phase(ctx, kDiag);
        if (st != ICQB_OK) return st;
    }
    int st = launch_offdiag(ctx, G, K, ld);
    if (st != ICQB_OK) return st;
    return finish_phase(ctx, kOffDiag);
}

int icqb_assemble_lower_dev(icqb_ctx* ctx, int diag_order, uint64_t dG, uint64_t dKl, uint64_t dKu) {
    if (!ctx) return ICQB_EARG;
    if (ctx->ntri == 0) return ICQB_OK;
    if (!ctx->d_qp_aos) return set_error(ctx, ICQB_ESTATE, "icqb_assemble_lower_dev: no mesh set");
    if (!dG) return set_error(ctx, ICQB_EARG, "icqb_assemble_lower_dev: null G");
    if ((dKl == 0) != (dKu == 0))
        return set_error(ctx, ICQB_EARG, "icqb_assemble_lower_dev: pass both K buffers or neither");
    if (diag_order != 0 && (diag_order < 1 || diag_order > 5))
        return set_error(ctx, ICQB_EORDER,
                         "diagonal order must be in 1..5 (KeyError in the reference)");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    double* G = reinterpret_cast<double*>(dG);
    LowerLayout L;
    int st = lower_layout(ctx, &L);
    if (st != ICQB_OK) return st;
    const long long rows = (L.r1 - L.r0) + (L.q1 - L.q0);
    if (diag_order != 0 && rows > 0) {
        // self terms into a compact vector, then placed on the diagonal of the diagonal tiles
        st = ensure_work(ctx, rows + 64);
        if (st != ICQB_OK) return st;
        cudaEventRecord(ctx->ev0, ctx->stream);
        st = launch_diag_rows(ctx, diag_order, L.r0, L.r1, ctx->d_work, 1);
        if (st == ICQB_OK)
            st = launch_diag_rows(ctx, diag_order, L.q0, L.q1, ctx->d_work + (L.r1 - L.r0), 1);
        if (st != ICQB_OK) return st;
        k_scatter_diag_lower<<<128, 256, 0, ctx->stream>>>(ctx->d_work, L, G);
        ctx->launches += 1;
        cudaEventRecord(ctx->ev1, ctx->stream);
        st = finish_phase(ctx, kDiag);
        if (st != ICQB_OK) return st;
    }
    st = launch_offdiag_lower(ctx, G, reinterpret_cast<double*>(dKl), reinterpret_cast<double*>(dKu));
    if (st != ICQB_OK) return st;
    return finish_phase(ctx, kOffDiag);
}

int icqb_symv_dev(icqb_ctx* ctx, uint64_t dA, uint64_t dx, uint64_t dy, int scaled) {
    if (!ctx) return ICQB_EARG;
    if (!dA || !dx || !dy) return set_error(ctx, ICQB_EARG, "icqb_symv_dev: null pointer");
    if (!ctx->d_area2) return set_error(ctx, ICQB_ESTATE, "icqb_symv_dev: no mesh set");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->ntri;
    double* inv = nullptr;
    if (!scaled) {
        int st = ensure_work(ctx, n + 64);
        if (st != ICQB_OK) return st;
        inv = ctx->d_work;
        st = launch_reciprocal(ctx, ctx->d_area2, inv, n);
        if (st != ICQB_OK) return st;
    }
    cudaEventRecord(ctx->ev0, ctx->stream);
    int st = launch_symv(ctx, reinterpret_cast<const double*>(dA),
                         reinterpret_cast<const double*>(dx), ctx->d_area2,
                         scaled ? ctx->d_area2 : nullptr, scaled ? nullptr : inv,
                         reinterpret_cast<double*>(dy), ctx->stream);
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (st != ICQB_OK) return st;
    st = comm_allreduce_sum(ctx, reinterpret_cast<double*>(dy), n, ctx->stream);
    if (st != ICQB_OK) return st;
    return finish_phase(ctx, kGemv);
}

int icqb_symv_f32_dev(icqb_ctx* ctx, uint64_t dA, uint64_t dx, uint64_t dy, int scaled) {
    if (!ctx) return ICQB_EARG;
    if (!dA || !dx || !dy) return set_error(ctx, ICQB_EARG, "icqb_symv_f32_dev: null pointer");
    if (!ctx->d_area2) return set_error(ctx, ICQB_ESTATE, "icqb_symv_f32_dev: no mesh set");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->ntri;
    LowerLayout L;
    int st = lower_layout(ctx, &L);
    if (st != ICQB_OK) return st;
    const long long count = L.numTiles() * kLowerTileElems;
    if (count > ctx->a32_floats) {
        dev_free(ctx, ctx->d_A32);
        ctx->d_A32 = nullptr;
        ctx->a32_floats = 0;
        ICQB_CUDA(ctx, dev_alloc(ctx, &ctx->d_A32, sizeof(float) * (size_t)(count > 4 ? count : 4)));
        ctx->a32_floats = count;
    }
    st = launch_f64_to_f32(ctx, reinterpret_cast<const double*>(dA), ctx->d_A32, count, ctx->stream);
    if (st != ICQB_OK) return st;
    double* inv = nullptr;
    if (!scaled) {
        st = ensure_work(ctx, n + 64);
        if (st != ICQB_OK) return st;
        inv = ctx->d_work;
        st = launch_reciprocal(ctx, ctx->d_area2, inv, n);
        if (st != ICQB_OK) return st;
    }
    cudaEventRecord(ctx->ev0, ctx->stream);
    st = launch_symv_f32_only(ctx, ctx->d_A32, reinterpret_cast<const double*>(dx), ctx->d_area2,
                              scaled ? ctx->d_area2 : nullptr, scaled ? nullptr : inv,
                              reinterpret_cast<double*>(dy), ctx->stream);
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (st != ICQB_OK) return st;
    st = comm_allreduce_sum(ctx, reinterpret_cast<double*>(dy), n, ctx->stream);
    if (st != ICQB_OK) return st;
    return finish_phase(ctx, kGemv);
}

int icqb_assemble_host(icqb_ctx* ctx, int diag_order, double* gMat, double* kMat) {
    if (!ctx || !gMat) return set_error(ctx, ICQB_EARG, "icqb_assemble_host: null output");
    const int64_t rows = ctx->r1 - ctx->r0, n = ctx->ntri;
    if (rows <= 0 || n == 0) return ICQB_OK;
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dG = nullptr, *dK = nullptr;
    ICQB_CUDA(ctx, cudaMalloc(&dG, sizeof(double) * rows * n));
    int st = ICQB_OK;
    if (kMat) st = check_cuda(ctx, cudaMalloc(&dK, sizeof(double) * rows * n), "cudaMalloc K");
    // diag_order == 0 keeps the caller's diagonal: stage it in
    if (st == ICQB_OK && diag_order == 0)
        st = check_cuda(ctx,
                        cudaMemcpy2DAsync(dG + ctx->r0, sizeof(double) * (n + 1), gMat + ctx->r0,
                                          sizeof(double) * (n + 1), sizeof(double), rows,
                                          cudaMemcpyHostToDevice, ctx->stream),
                        "diag upload");
    if (st == ICQB_OK) st = icqb_assemble_dev(ctx, diag_order, (uint64_t)dG, (uint64_t)dK, n);
    if (st == ICQB_OK)
        st = check_cuda(ctx, cudaMemcpy(gMat, dG, sizeof(double) * rows * n, cudaMemcpyDeviceToHost),
                        "G download");
    if (st == ICQB_OK && kMat)
        st = check_cuda(ctx, cudaMemcpy(kMat, dK, sizeof(double) * rows * n, cudaMemcpyDeviceToHost),
                        "K download");
    cudaFree(dG);
    cudaFree(dK);
    return st;
}

int icqb_diagonal_host(icqb_ctx* ctx, int diag_order, double* diag) {
    if (!ctx || !diag) return set_error(ctx, ICQB_EARG, "icqb_diagonal_host: null output");
    const int64_t rows = ctx->r1 - ctx->r0;
    if (rows <= 0) return ICQB_OK;
    if (!ctx->d_verts) return set_error(ctx, ICQB_ESTATE, "icqb_diagonal_host: no mesh set");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;
    ICQB_CUDA(ctx, cudaMalloc(&d, sizeof(double) * rows));
    int st = launch_diag(ctx, diag_order, d, 1);
    if (st == ICQB_OK) st = finish_phase(ctx, kDiag);
    if (st == ICQB_OK)
        st = check_cuda(ctx, cudaMemcpy(diag, d, sizeof(double) * rows, cudaMemcpyDeviceToHost),
                        "diag download");
    cudaFree(d);
    return st;
}

int computeOffDiagonalTermsArrays(const double* xyz, int64_t npts, const int64_t* tri,
                                  int64_t ntri, double* gMat) {
    icqb_ctx* ctx = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    int st = icqb_create(&ctx, dev);
    if (st != ICQB_OK) return st;
    st = icqb_set_mesh(ctx, xyz, npts, tri, ntri);
    if (st == ICQB_OK) st = icqb_assemble_host(ctx, 0, gMat, nullptr);
    if (st != ICQB_OK) set_error(nullptr, st, ctx->err);
    icqb_destroy(ctx);
    return st;
}

int icqb_area2_host(icqb_ctx* ctx, double* out) {
    if (!ctx || !out) return ICQB_EARG;
    if (ctx->ntri == 0) return ICQB_OK;
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ICQB_CUDA(ctx, cudaMemcpy(out, ctx->d_area2, sizeof(double) * ctx->ntri, cudaMemcpyDeviceToHost));
    return ICQB_OK;
}

int icqb_gemv_dev(icqb_ctx* ctx, uint64_t dA, int64_t nrows, int64_t ncols, int64_t ld,
                  uint64_t dx, uint64_t dy, uint64_t drowscale, uint64_t stream) {
    if (!ctx) return ICQB_EARG;
    if (!dA || !dx || !dy) return set_error(ctx, ICQB_EARG, "icqb_gemv_dev: null pointer");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : ctx->stream;
    return launch_gemv(ctx, reinterpret_cast<const double*>(dA), nrows, ncols, ld,
                       reinterpret_cast<const double*>(dx), reinterpret_cast<double*>(dy),
                       reinterpret_cast<const double*>(drowscale), st);
}

int icqb_gemv_host(icqb_ctx* ctx, uint64_t dA, int64_t nrows, int64_t ncols, int64_t ld,
                   const double* x, double* y) {
    if (!ctx || !x || !y) return set_error(ctx, ICQB_EARG, "icqb_gemv_host: null pointer");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    int st = ensure_work(ctx, ncols + nrows + 2);
    if (st != ICQB_OK) return st;
    double* dx = ctx->d_work;
    double* dy = ctx->d_work + ((ncols + 1) & ~1LL);
    ICQB_CUDA(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * ncols, cudaMemcpyHostToDevice, ctx->stream));
    cudaEventRecord(ctx->ev0, ctx->stream);
    st = launch_gemv(ctx, reinterpret_cast<const double*>(dA), nrows, ncols, ld, dx, dy, nullptr,
                     ctx->stream);
    cudaEventRecord(ctx->ev1, ctx->stream);
    if (st != ICQB_OK) return st;
    ICQB_CUDA(ctx, cudaMemcpyAsync(y, dy, sizeof(double) * nrows, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_phase(ctx, kGemv);
}

int icqb_measure_fp64_peak(icqb_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return ICQB_EARG;
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 1 << 15;
    int st = ensure_work(ctx, (int64_t)blocks * threads);
    if (st != ICQB_OK) return st;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(ctx->d_work, iters);
        cudaEventRecord(ctx->ev1, ctx->stream);
        ctx->launches += 1;
        ICQB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        const double flop = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) best = flop / (ms * 1e-3) / 1e12 > best ? flop / (ms * 1e-3) / 1e12 : best;
    }
    *tflops = best;
    return ICQB_OK;
}

}  // extern "C"
