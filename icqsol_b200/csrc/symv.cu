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
// Mapping: one CTA (8 warps) per task = up to 8 consecutive tiles of one 128-row
// strip.  A warp owns 16 rows of the strip; its 32 lanes cover the 128 columns of
// a tile as 4 consecutive doubles each (one 1 KB contiguous segment per row, two
// 16-byte streaming loads per lane, four rows in flight).  Row sums stay in
// registers for the whole task and are shuffled down once at its end; column sums
// are combined across the 8 warps through shared memory once per tile and stored
// to a per-strip workspace.  A second small kernel adds the partials in a fixed
// order (no atomics: results are reproducible) and applies alpha/beta.
#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"

namespace icqb {

namespace {

constexpr int kT = 128;          // tile edge
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRowsPerWarp = kT / kWarps;   // 16
constexpr int kChunk = 4;        // tiles per task (512 KB of matrix)

struct SymvTask {
    int lt;        // local strip (tile row) index
    int j0, j1;    // global source tiles [j0, j1), j1 - 1 <= I
};

struct SymvParams {
    const double* A;
    long long ld, n;
    long long r0, r1, q0, q1;   // row blocks (multiples of 128 at the start)
    int nTiles0, nLocal;        // strips of the first block / all local strips
    const SymvTask* tasks;
    const int* order;           // launch order: longest tasks first
    const int* taskBegin;       // [nLocal+1] first task of each strip
    const double* x;            // [n]
    const double* gamma;        // [n] or null
    const double* alpha;        // [n] or null
    const double* beta;         // [n] or null
    double* Wc;                 // [nLocal][wcols] column partials per strip
    double* Wr;                 // [nTasks][128]  row partials per task
    long long wcols;
    double* z;                  // [n]
    const int* skip;            // when non-null and *skip != 0 the kernels return at once
};

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ void strip_of(const SymvParams& P, int lt, long long* obs0,
                                         long long* obsEnd, long long* lrow0) {
    if (lt < P.nTiles0) {
        *obs0 = P.r0 + (long long)lt * kT;
        *obsEnd = P.r1;
        *lrow0 = (long long)lt * kT;
    } else {
        *obs0 = P.q0 + (long long)(lt - P.nTiles0) * kT;
        *obsEnd = P.q1;
        *lrow0 = (P.r1 - P.r0) + (long long)(lt - P.nTiles0) * kT;
    }
}

__global__ void __launch_bounds__(kThreads, 2) k_symv(const SymvParams P) {
    __shared__ double s_u[kT];                    // gamma . x for the strip's rows
    __shared__ double s_col[2][kWarps][kT];       // column partials, double buffered

    if (P.skip && *P.skip) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int taskId = P.order[blockIdx.x];
    const SymvTask task = P.tasks[taskId];
    long long obs0, obsEnd, lrow0;
    strip_of(P, task.lt, &obs0, &obsEnd, &lrow0);
    const int I = (int)(obs0 / kT);
    const int nrows = (int)min((long long)kT, obsEnd - obs0);

    if (tid < kT) {
        double u = 0.0;
        if (tid < nrows) {
            const long long i = obs0 + tid;
            u = P.x[i];
            if (P.gamma) u *= P.gamma[i];
        }
        s_u[tid] = u;
    }
    __syncthreads();

    double rowacc[kRowsPerWarp];
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) rowacc[k] = 0.0;
    const double* aBase = P.A + lrow0 * P.ld;

    int buf = 0;
    for (int J = task.j0; J < task.j1; ++J) {
        const bool diagTile = (J == I);
        const long long c0 = (long long)J * kT + lane * 4;
        // valid columns of this tile: everything below the diagonal tile is complete;
        // the diagonal tile stops at n
        const long long cEnd = diagTile ? min(P.n, (long long)(J + 1) * kT) : (long long)(J + 1) * kT;
        const bool full4 = (c0 + 4 <= cEnd);
        double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
        if (full4) {
            x0 = __ldg(P.x + c0); x1 = __ldg(P.x + c0 + 1);
            x2 = __ldg(P.x + c0 + 2); x3 = __ldg(P.x + c0 + 3);
        } else {
            if (c0 < cEnd) x0 = __ldg(P.x + c0);
            if (c0 + 1 < cEnd) x1 = __ldg(P.x + c0 + 1);
            if (c0 + 2 < cEnd) x2 = __ldg(P.x + c0 + 2);
        }
        double ca0 = 0, ca1 = 0, ca2 = 0, ca3 = 0;

#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; rr += 4) {
            double2 lo[4], hi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = warp + kWarps * (rr + u);
                const double* ap = aBase + (long long)r * P.ld + c0;
                lo[u] = make_double2(0.0, 0.0);
                hi[u] = make_double2(0.0, 0.0);
                if (r < nrows) {
                    if (full4) {
                        lo[u] = ld_stream2(ap);
                        hi[u] = ld_stream2(ap + 2);
                    } else {
                        if (c0 < cEnd) lo[u].x = __ldg(ap);
                        if (c0 + 1 < cEnd) lo[u].y = __ldg(ap + 1);
                        if (c0 + 2 < cEnd) hi[u].x = __ldg(ap + 2);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = warp + kWarps * (rr + u);
                rowacc[rr + u] = fma(lo[u].x, x0, rowacc[rr + u]);
                rowacc[rr + u] = fma(lo[u].y, x1, rowacc[rr + u]);
                rowacc[rr + u] = fma(hi[u].x, x2, rowacc[rr + u]);
                rowacc[rr + u] = fma(hi[u].y, x3, rowacc[rr + u]);
                if (!diagTile) {
                    const double ur = s_u[r];
                    ca0 = fma(lo[u].x, ur, ca0);
                    ca1 = fma(lo[u].y, ur, ca1);
                    ca2 = fma(hi[u].x, ur, ca2);
                    ca3 = fma(hi[u].y, ur, ca3);
                }
            }
        }

        if (!diagTile) {
            double* sc = &s_col[buf][warp][lane * 4];
            sc[0] = ca0; sc[1] = ca1; sc[2] = ca2; sc[3] = ca3;
            __syncthreads();
            if (tid < kT) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) v += s_col[buf][w][tid];
                P.Wc[(long long)task.lt * P.wcols + (long long)J * kT + tid] = v;
            }
            buf ^= 1;   // the next tile writes the other buffer; its barrier orders the reuse
        }
    }

    // row partials of this task: one shuffle reduction per row, fixed order
#pragma unroll
    for (int k = 0; k < kRowsPerWarp; ++k) {
        double v = rowacc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) P.Wr[(long long)taskId * kT + warp + kWarps * k] = v;
    }
}

// z[j] = alpha_j * (sum of row partials of j's strip) + beta_j * (sum over strips below of Wc)
__global__ void __launch_bounds__(256) k_symv_reduce(const SymvParams P) {
    if (P.skip && *P.skip) return;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.n) return;
    const int J = (int)(j / kT);
    double col = 0.0;
    for (int lt = 0; lt < P.nLocal; ++lt) {
        const long long obs0 = (lt < P.nTiles0) ? P.r0 + (long long)lt * kT
                                                 : P.q0 + (long long)(lt - P.nTiles0) * kT;
        if ((int)(obs0 / kT) > J) col += P.Wc[(long long)lt * P.wcols + j];
    }
    double row = 0.0;
    int lt = -1;
    if (j >= P.r0 && j < P.r1) lt = (int)((j - P.r0) / kT);
    else if (j >= P.q0 && j < P.q1) lt = P.nTiles0 + (int)((j - P.q0) / kT);
    if (lt >= 0) {
        const int within = (int)(j % kT);
        for (int t = P.taskBegin[lt]; t < P.taskBegin[lt + 1]; ++t)
            row += P.Wr[(long long)t * kT + within];
    }
    if (P.alpha) row *= P.alpha[j];
    if (P.beta) col *= P.beta[j];
    P.z[j] = row + col;
}

}  // namespace

// cached task table / workspace of a context, rebuilt when the row blocks change
struct SymvPlan {
    long long n = -1, r0 = 0, r1 = 0, q0 = 0, q1 = 0;
    int nTiles0 = 0, nLocal = 0, nTasks = 0;
    long long wcols = 0;
    SymvTask* d_tasks = nullptr;
    int* d_order = nullptr;
    int* d_taskBegin = nullptr;
    double* d_Wc = nullptr;
    double* d_Wr = nullptr;
};

void symv_plan_free(icqb_ctx* ctx) {
    SymvPlan* p = ctx->symv;
    if (!p) return;
    cudaFree(p->d_tasks);
    cudaFree(p->d_order);
    cudaFree(p->d_taskBegin);
    cudaFree(p->d_Wc);
    cudaFree(p->d_Wr);
    delete p;
    ctx->symv = nullptr;
}

static int symv_plan(icqb_ctx* ctx) {
    SymvPlan* p = ctx->symv;
    if (p && p->n == ctx->ntri && p->r0 == ctx->r0 && p->r1 == ctx->r1 && p->q0 == ctx->q0 &&
        p->q1 == ctx->q1)
        return ICQB_OK;
    symv_plan_free(ctx);
    if ((ctx->r1 > ctx->r0 && (ctx->r0 % kT) != 0) || (ctx->q1 > ctx->q0 && (ctx->q0 % kT) != 0))
        return set_error(ctx, ICQB_EARG, "lower storage: row blocks must start on multiples of 128");
    p = new SymvPlan();
    ctx->symv = p;
    p->n = ctx->ntri;
    p->r0 = ctx->r0; p->r1 = ctx->r1; p->q0 = ctx->q0; p->q1 = ctx->q1;
    p->nTiles0 = (int)((p->r1 - p->r0 + kT - 1) / kT);
    const int nTiles1 = (int)((p->q1 - p->q0 + kT - 1) / kT);
    p->nLocal = p->nTiles0 + nTiles1;
    p->wcols = (p->n + kT - 1) / kT * kT;
    std::vector<SymvTask> tasks;
    std::vector<int> begin(p->nLocal + 1, 0);
    for (int lt = 0; lt < p->nLocal; ++lt) {
        const long long obs0 = lt < p->nTiles0 ? p->r0 + (long long)lt * kT
                                               : p->q0 + (long long)(lt - p->nTiles0) * kT;
        const int I = (int)(obs0 / kT);
        begin[lt] = (int)tasks.size();
        for (int j0 = 0; j0 <= I; j0 += kChunk)
            tasks.push_back(SymvTask{lt, j0, std::min(I + 1, j0 + kChunk)});
    }
    begin[p->nLocal] = (int)tasks.size();
    p->nTasks = (int)tasks.size();
    if (p->nTasks == 0) return ICQB_OK;
    std::vector<int> order(tasks.size());
    for (size_t t = 0; t < order.size(); ++t) order[t] = (int)t;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (tasks[a].j1 - tasks[a].j0) > (tasks[b].j1 - tasks[b].j0);
    });
    ICQB_CUDA(ctx, cudaMalloc(&p->d_order, sizeof(int) * order.size()));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_order, order.data(), sizeof(int) * order.size(),
                              cudaMemcpyHostToDevice));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_tasks, sizeof(SymvTask) * tasks.size()));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_taskBegin, sizeof(int) * begin.size()));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_Wc, sizeof(double) * (size_t)p->nLocal * p->wcols));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_Wr, sizeof(double) * (size_t)p->nTasks * kT));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_tasks, tasks.data(), sizeof(SymvTask) * tasks.size(),
                              cudaMemcpyHostToDevice));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_taskBegin, begin.data(), sizeof(int) * begin.size(),
                              cudaMemcpyHostToDevice));
    return ICQB_OK;
}

int launch_symv(icqb_ctx* ctx, const double* dA, int64_t ld, const double* dx,
                const double* dgamma, const double* dalpha, const double* dbeta, double* dz,
                cudaStream_t stream, const int* skip) {
    if (ctx->ntri == 0) return ICQB_OK;
    if ((reinterpret_cast<uintptr_t>(dA) & 15) != 0 || (ld & 1) != 0 || ld < ctx->ntri)
        return set_error(ctx, ICQB_EARG, "symv: matrix must be 16-byte aligned with an even ld >= n");
    int st = symv_plan(ctx);
    if (st != ICQB_OK) return st;
    SymvPlan* p = ctx->symv;
    SymvParams P;
    P.A = dA;
    P.ld = ld;
    P.n = p->n;
    P.r0 = p->r0; P.r1 = p->r1; P.q0 = p->q0; P.q1 = p->q1;
    P.nTiles0 = p->nTiles0;
    P.nLocal = p->nLocal;
    P.tasks = p->d_tasks;
    P.order = p->d_order;
    P.taskBegin = p->d_taskBegin;
    P.x = dx;
    P.gamma = dgamma;
    P.alpha = dalpha;
    P.beta = dbeta;
    P.Wc = p->d_Wc;
    P.Wr = p->d_Wr;
    P.wcols = p->wcols;
    P.z = dz;
    P.skip = skip;
    if (p->nTasks > 0) {
        k_symv<<<p->nTasks, kThreads, 0, stream>>>(P);
        ctx->launches += 1;
    }
    k_symv_reduce<<<(unsigned)((p->n + 255) / 256), 256, 0, stream>>>(P);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_symv launch");
}

}  // namespace icqb
