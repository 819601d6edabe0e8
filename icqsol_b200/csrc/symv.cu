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
// Mapping: one WARP per 128x128 tile, four independent warps per CTA, no shared
// memory and no barriers.  The 32 lanes cover the tile's 128 columns as 4
// consecutive doubles each (a 1 KB contiguous segment per row, two 16-byte streaming
// loads per lane) and walk the 128 rows eight at a time (16 loads in flight per
// lane, 128 KB per SM at 16 resident warps).  Column sums are complete inside the
// warp (each lane owns its 4 columns over all rows); a row's partial is reduced over
// the lanes by a shuffle butterfly and kept by lane (row mod 32).  Each tile stores
// 128 row partials and 128 column partials (2 KB per 128 KB streamed); a second
// kernel adds them per global index in a fixed order (no atomics: results are
// reproducible) and applies alpha/beta.
#include <algorithm>
#include <vector>

#include "icqb_internal.cuh"

namespace icqb {

namespace {

constexpr int kT = 128;          // tile edge
constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kRowsInFlight = 8;

struct SymvTask {
    int lt;        // local strip (tile row) index
    int J;         // global source tile, J <= I
};

struct SymvParams {
    const double* A;            // tile-contiguous lower storage (LowerLayout)
    long long n;
    LowerLayout L;
    int nTasks;                 // == number of stored tiles; task t works on tile t
    const SymvTask* tasks;
    const double* x;            // [n]
    const double* gamma;        // [n] or null
    const double* alpha;        // [n] or null
    const double* beta;         // [n] or null
    double* Wc;                 // [nLocal][wcols] column partials per strip
    double* Wr;                 // [nTasks][128]  row partials per tile
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

__global__ void __launch_bounds__(kThreads, 4) k_symv(const SymvParams P) {
    if (P.skip && *P.skip) return;
    const int lane = threadIdx.x & 31;
    const int taskId = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (taskId >= P.nTasks) return;
    const SymvTask task = P.tasks[taskId];
    const long long obs0 = P.L.obs0(task.lt), obsEnd = P.L.obsEnd(task.lt);
    const int I = (int)(obs0 / kT), J = task.J;
    const int nrows = (int)min((long long)kT, obsEnd - obs0);
    const bool diagTile = (J == I);

    const long long c0 = (long long)J * kT + lane * 4;
    // valid columns: tiles below the diagonal tile are complete; the diagonal tile ends at n
    const long long cEnd = diagTile ? min(P.n, (long long)(J + 1) * kT) : (long long)(J + 1) * kT;
    const bool full4 = (c0 + 4 <= cEnd);
    double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    if (c0 < cEnd) x0 = __ldg(P.x + c0);
    if (c0 + 1 < cEnd) x1 = __ldg(P.x + c0 + 1);
    if (c0 + 2 < cEnd) x2 = __ldg(P.x + c0 + 2);
    if (c0 + 3 < cEnd) x3 = __ldg(P.x + c0 + 3);

    double ca0 = 0, ca1 = 0, ca2 = 0, ca3 = 0;      // column sums of this lane's 4 columns
    double rowacc[kT / 32] = {0, 0, 0, 0};          // rows lane, lane+32, lane+64, lane+96
    // the tile is one contiguous 128 KB block: row r starts at r*128
    const double* aBase = P.A + (long long)taskId * kLowerTileElems + lane * 4;

#pragma unroll
    for (int q = 0; q < kT / 32; ++q) {
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += kRowsInFlight) {
            const int rbase = q * 32 + rb;
            if (rbase >= nrows) break;   // uniform across the warp
            double2 lo[kRowsInFlight], hi[kRowsInFlight];
#pragma unroll
            for (int u = 0; u < kRowsInFlight; ++u) {
                const int r = rbase + u;
                const double* ap = aBase + r * kT;
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
            for (int u = 0; u < kRowsInFlight; ++u) {
                const int r = rbase + u;
                double part = lo[u].x * x0;
                part = fma(lo[u].y, x1, part);
                part = fma(hi[u].x, x2, part);
                part = fma(hi[u].y, x3, part);
                if (!diagTile && r < nrows) {
                    const long long i = obs0 + r;
                    double ur = __ldg(P.x + i);            // uniform address: broadcast
                    if (P.gamma) ur *= __ldg(P.gamma + i);
                    ca0 = fma(lo[u].x, ur, ca0);
                    ca1 = fma(lo[u].y, ur, ca1);
                    ca2 = fma(hi[u].x, ur, ca2);
                    ca3 = fma(hi[u].y, ur, ca3);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                    part += __shfl_xor_sync(0xffffffffu, part, off);
                if (lane == rb + u) rowacc[q] = part;      // row (q*32 + lane) of the tile
            }
        }
    }

    double* wr = P.Wr + (long long)taskId * kT;
#pragma unroll
    for (int q = 0; q < kT / 32; ++q) wr[q * 32 + lane] = rowacc[q];
    if (!diagTile) {
        double* wc = P.Wc + (long long)task.lt * P.wcols + c0;
        reinterpret_cast<double2*>(wc)[0] = make_double2(ca0, ca1);
        reinterpret_cast<double2*>(wc)[1] = make_double2(ca2, ca3);
    }
}

// z[j] = alpha_j * (row partials of j's strip over its tiles) + beta_j * (column partials of
// the strips below).  One CTA per 128 global indices, 4 thread groups split the two sums and
// are combined in a fixed order.
__global__ void __launch_bounds__(512) k_symv_reduce(const SymvParams P) {
    __shared__ double s_row[4][kT], s_col[4][kT];
    if (P.skip && *P.skip) return;
    const int T = blockIdx.x, c = threadIdx.x, g = threadIdx.y;
    const long long j = (long long)T * kT + c;
    const long long jl = min(j, P.n - 1);       // clamp: lanes beyond n never store
    double col = 0.0, row = 0.0;
    // strips strictly below tile row T, first block then second block
    {
        const long long first = max(0LL, (long long)T + 1 - P.L.r0 / kT);
        for (long long lt = first + g; lt < P.L.nTiles0; lt += 4) col += P.Wc[lt * P.wcols + jl];
        const long long first1 = max(0LL, (long long)T + 1 - P.L.q0 / kT);
        for (long long lt = P.L.nTiles0 + first1 + g; lt < P.L.nLocal(); lt += 4)
            col += P.Wc[lt * P.wcols + jl];
    }
    // this tile row, if it is stored here
    const int lt = P.L.stripOf((long long)T * kT);
    if (lt >= 0) {
        const long long base = P.L.tileBase(lt);
        for (int Jt = g; Jt <= T; Jt += 4) row += P.Wr[(base + Jt) * kT + c];
    }
    s_row[g][c] = row;
    s_col[g][c] = col;
    __syncthreads();
    if (g == 0 && j < P.n) {
        row = (s_row[0][c] + s_row[1][c]) + (s_row[2][c] + s_row[3][c]);
        col = (s_col[0][c] + s_col[1][c]) + (s_col[2][c] + s_col[3][c]);
        if (P.alpha) row *= P.alpha[j];
        if (P.beta) col *= P.beta[j];
        P.z[j] = row + col;
    }
}

}  // namespace

// cached task table / workspace of a context, rebuilt when the row blocks change
struct SymvPlan {
    long long n = -1;
    LowerLayout L{0, 0, 0, 0, 0, 0, 0};
    int nTasks = 0;
    long long wcols = 0;
    SymvTask* d_tasks = nullptr;
    double* d_Wc = nullptr;
    double* d_Wr = nullptr;
};

void symv_plan_free(icqb_ctx* ctx) {
    SymvPlan* p = ctx->symv;
    if (!p) return;
    cudaFree(p->d_tasks);
    cudaFree(p->d_Wc);
    cudaFree(p->d_Wr);
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
    std::vector<SymvTask> tasks;
    for (int lt = 0; lt < L.nLocal(); ++lt) {
        const int I = (int)L.tileRow(lt);
        for (int J = 0; J <= I; ++J) tasks.push_back(SymvTask{lt, J});
    }
    p->nTasks = (int)tasks.size();
    if ((long long)p->nTasks != L.numTiles())
        return set_error(ctx, ICQB_ESTATE, "symv: tile enumeration does not match the layout");
    if (p->nTasks == 0) return ICQB_OK;
    ICQB_CUDA(ctx, cudaMalloc(&p->d_tasks, sizeof(SymvTask) * tasks.size()));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_Wc, sizeof(double) * (size_t)L.nLocal() * p->wcols));
    ICQB_CUDA(ctx, cudaMalloc(&p->d_Wr, sizeof(double) * (size_t)p->nTasks * kT));
    ICQB_CUDA(ctx, cudaMemcpy(p->d_tasks, tasks.data(), sizeof(SymvTask) * tasks.size(),
                              cudaMemcpyHostToDevice));
    return ICQB_OK;
}

int launch_symv(icqb_ctx* ctx, const double* dA, const double* dx, const double* dgamma,
                const double* dalpha, const double* dbeta, double* dz, cudaStream_t stream,
                const int* skip) {
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
    P.wcols = p->wcols;
    P.z = dz;
    P.skip = skip;
    if (p->nTasks > 0) {
        k_symv<<<(p->nTasks + kWarpsPerCta - 1) / kWarpsPerCta, kThreads, 0, stream>>>(P);
        ctx->launches += 1;
    }
    k_symv_reduce<<<(unsigned)((p->n + kT - 1) / kT), dim3(kT, 4), 0, stream>>>(P);
    ctx->launches += 1;
    return check_cuda(ctx, cudaGetLastError(), "k_symv launch");
}

}  // namespace icqb
