// The solver behind icqb_cg_solve_dev for preconditioner kind 2 (the default of the solver
// classes; validated on B200 in round 2 up to 199,808 triangles on 8 GPUs): conjugate gradient
// with DENSE DIAGONAL BLOCKS of several 128-row tiles in the block preconditioner.
//
// Why: (1) neighbouring sliver triangles (the rings at the poles of a UV sphere) make the
// reference's diag(A) G indefinite -- the wrong-sign eigenvectors live entirely on those
// rings -- and the CG of cg.cu with 128 x 128 blocks then does not
// converge; with ONE dense block over every run of sliver tiles (on a UV sphere the leading
// and the trailing tiles, "caps") it converges: 171 iterations at 100,488 triangles, 273 at
// 199,808 (tools/experiments/pcg_cap_blocks.py for the model).  (2) On regular meshes larger
// blocks simply pay: 91 iterations instead of 165 at config C2 with blocks of 8 tiles
// (tools/experiments/pcg_block_size.py).
//
// Same algorithm, state and product kernels as cg.cu (solvers/icqConjugateGradient.py:49-79);
// what differs is the preconditioner step w = M^-1 r:
//   * the tile rows are partitioned into diagonal blocks (icqb_cg_set_block_partition, or two
//     caps around single tiles, icqb_cg_set_cap_tiles); every block -- a single tile included --
//     is extracted from the stored tiles as a dense
//     matrix of diag(s) A (lower storage: symmetric fill) by the rank that holds its rows (or
//     all-reduced when it straddles ranks), inverted in place by a symmetric block sweep (below;
//     all blocks batched in a launch) -- once per solve;
//   * per iteration the residual kernel is split in two so that the dense blocks can be
//     applied in between:
//         k_cg_vec       x += alpha p, r -= alpha z            (all rows)
//         k_dense_apply  w = C^-1 r on the rows of the dense blocks (one warp per row)
//         k_cg_close     w = M^-1 r on the single tiles, rho, norms, beta, stopping flag.
//
// The same solver carries the optional PEER EXCHANGE (several ranks, lower storage, when
// icqb_peer_create/attach have been called): k_symv_reduce<PEER> stores the partial product
// and its p.z into every rank's exchange buffer and raises a flag there; k_peer_gather waits
// for the P flags of the product's sequence number and adds the P slots in rank order -- no
// NCCL call inside the iteration.  Slots are double-buffered by the parity of the sequence
// number; a rank can be at most one product ahead of a peer because it needs that peer's flag
// of the previous product, and restarts of the iteration pass through an NCCL all-reduce.
#include "cg_kernels.cuh"

namespace icqb {

namespace {

// z = sum over ranks of the partial products in this rank's exchange buffer (rank order:
// identical on every rank), p.z likewise; waits for every rank's flag first.
__global__ void __launch_bounds__(kT) k_peer_gather(const SymvPeer X, long long n, double* z,
                                                    double* pz, const int* skip) {
    __shared__ int s_ready;
    if (skip && *skip) return;
    const volatile unsigned long long* flags =
        reinterpret_cast<const volatile unsigned long long*>(X.flags[X.rank]) + X.parity * X.nranks;
    if (threadIdx.x == 0) {
        int ready = 1;
#ifdef ICQB_EMU
        for (int h = 0; h < X.nranks; ++h)
            while (flags[h] < X.seq) emu::yield();
#else
        // bounded wait (~20 s): a peer that died must not hang this GPU -- the product then
        // comes back as NaN and the solve stops with a breakdown
        const long long t0 = clock64();
        for (int h = 0; h < X.nranks && ready; ++h)
            while (flags[h] < X.seq)
                if (clock64() - t0 > 40000000000LL) {
                    ready = 0;
                    break;
                }
#endif
        __threadfence_system();
        s_ready = ready;
    }
    __syncthreads();
    if (!s_ready) {
        const long long i0 = (long long)blockIdx.x * kT + threadIdx.x;
        if (i0 < n) z[i0] = nan("");
        if (blockIdx.x == 0 && threadIdx.x == 0) *pz = nan("");
        return;
    }
    const double* slots = X.slots[X.rank] + (long long)X.parity * X.nranks * X.slotStride;
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) {
        double acc = 0.0;
        for (int h = 0; h < X.nranks; ++h)
            acc += reinterpret_cast<const volatile double*>(slots + h * X.slotStride)[i];
        z[i] = acc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int h = 0; h < X.nranks; ++h)
            acc += reinterpret_cast<const volatile double*>(slots + h * X.slotStride)[n];
        *pz = acc;
    }
}

// One dense diagonal block of the preconditioner: tile rows [t0, t0 + tiles), tiles > 1.
struct DenseDesc {
    long long t0, tiles;
    long long m;       // 128 * tiles: order of the block
    long long valid;   // rows of the block below n
    long long off;     // offset (doubles) of its m x m matrix in the dense buffer
    long long row0;    // first compact row (valid rows of all dense blocks, concatenated)
    long long row0pad; // sum of m over the blocks before this one (position in the panel buffers)
    // overlapping blocks (one rank): t0 / tiles / m / valid describe the EXTENT; the block proper
    // is the global rows [core0, core1); urow0 = first row of the extent in the sequence of all
    // extents, each padded to a multiple of the update kernel's rows per CTA
    long long core0, core1, urow0;
};

// dense block of diag(scale) A from the lower storage: tiles J <= I are copied (rows this rank
// stores), tiles J < I also mirrored -- diag(s) A is symmetric.  The buffer is zeroed before;
// identity beyond n.  Grid (kmax, kmax, blocks).
__global__ void __launch_bounds__(256) k_dense_extract_lower(const double* A, LowerLayout L,
                                                             const double* scale, long long n,
                                                             const DenseDesc* desc, double* dense) {
    const DenseDesc D = desc[blockIdx.z];
    if (blockIdx.x >= D.tiles || blockIdx.y >= D.tiles) return;
    const long long I = D.t0 + blockIdx.y, J = D.t0 + blockIdx.x;
    if (J > I) return;
    const int lt = L.stripOf(I * kT);
    if (lt < 0) return;
    const double* a = A + (L.tileBase(lt) + J) * kLowerTileElems;
    double* out = dense + D.off;
    const long long m = D.m;
    for (int e = threadIdx.x; e < kT * kT; e += blockDim.x) {
        const int r = e >> 7, c = e & (kT - 1);
        const long long i = I * kT + r, j = J * kT + c;
        double v;
        if (i < n && j < n)
            v = (scale ? scale[i] : 1.0) * a[e];
        else
            v = (I == J && r == c) ? 1.0 : 0.0;
        const long long lr = (I - D.t0) * kT + r, lc = (J - D.t0) * kT + c;
        out[lr * m + lc] = v;
        if (J < I) out[lc * m + lr] = v;
    }
}

// the same block from row-block storage (every tile of the block, rows this rank holds)
__global__ void __launch_bounds__(256) k_dense_extract_full(const double* A, long long ld,
                                                            long long r0, long long r1,
                                                            const double* scale, long long n,
                                                            const DenseDesc* desc, double* dense) {
    const DenseDesc D = desc[blockIdx.z];
    if (blockIdx.x >= D.tiles || blockIdx.y >= D.tiles) return;
    const long long I = D.t0 + blockIdx.y, J = D.t0 + blockIdx.x;
    const bool ownsLast = (n - 1 >= r0 && n - 1 < r1);
    double* out = dense + D.off;
    for (int e = threadIdx.x; e < kT * kT; e += blockDim.x) {
        const int r = e >> 7, c = e & (kT - 1);
        const long long i = I * kT + r, j = J * kT + c;
        double v = 0.0;
        if (i < n) {
            if (i >= r0 && i < r1 && j < n) v = (scale ? scale[i] : 1.0) * A[(i - r0) * ld + j];
        } else if (ownsLast) {
            v = (I == J && r == c) ? 1.0 : 0.0;
        }
        out[((I - D.t0) * kT + r) * D.m + (J - D.t0) * kT + c] = v;
    }
}

// In-place inverse of every dense block (symmetric, m = 128 tiles): BLOCK SWEEP on the lower
// block triangle, no pivoting.  Gauss-Jordan on a symmetric matrix keeps, after the pivot tiles
// S = {0..K-1} have been swept, [S,S] = A_SS^-1 and [U,U] = the Schur complement symmetric and
// [S,U] = -[U,S]^T, so the upper block triangle never has to be formed.  Step K, with
// V_J (128 x 128, k-major) = A_KJ for J < K and A_JK^T for J > K ("row K of the completed
// matrix"):
//   (a) k_sweep_pivot    P = A_KK^-1 in registers (tile_gauss_jordan), symmetrised, in place
//   (b) k_sweep_panel    W_J = P V_J for every J != K; V and W go to two 128 x m panel buffers
//   (c) k_sweep_update   A_IJ += s_I V_I^T W_J for I >= J, both != K; s_I = +1 (I < K), -1 (I > K)
//   (d) k_sweep_finish   row K: A_KJ = W_J (J < K); column K: A_IK = -W_I^T (I > K)
// i.e. (T - 1) + T (T - 1) / 2 tile products per step instead of the (T - 1) (T + 1) of the
// unsymmetric block Gauss-Jordan this replaces (T = 8: 35 instead of 63), and after the last step
// k_dense_mirror copies the lower triangle into the upper one (k_dense_apply streams full rows).
// Round 2 on a B200, config C2 (20 blocks of 8 tiles): the first version took 7.5 ms of a 40 ms
// solve -- 354 us per pivot tile inverted in shared memory, tile products at 9 TFLOP/s.
//
// (b) and (c) share one FP64 tile kernel: CTA tile 128 x 64, 128 threads, 8 x 8 outputs per
// thread (rows {2 tx, 2 tx + 1} + 32 i, columns 8 ty .. 8 ty + 7), K = 128 in slabs of 16 through
// double-buffered shared memory, the next slab's global loads in flight while the current one is
// multiplied (the arrangement of k_lu_gemm, lu.cu); both operands are k-major.
constexpr int kSwM = 128, kSwN = 64, kSwK = 16;
constexpr int kSwLdB = kSwN;       // (no padding: 2 x 16 x (128 + 64) doubles are the 48 KB of static shared memory)

__global__ void __launch_bounds__(kInvThreads) k_sweep_pivot(const DenseDesc* desc, double* dense,
                                                             long long K, int* weak, int* bad) {
    extern __shared__ double s_inv[];
    const DenseDesc D = desc[blockIdx.z];
    if (K >= D.tiles) return;
    double* s_t = s_inv;
    double* s_vec = s_inv + kT * kInvLd;
    double* g = dense + D.off + (K * kT) * D.m + K * kT;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;   // (tile_gauss_jordan begins with a barrier)
    double a[kGjRows][4];
    tile_to_regs(g, D.m, a);
    const int nweak = tile_gauss_jordan(a, s_vec);
    if (D.tiles == 1) {
        // a block of one tile: as k_cg_invert_blocks -- a weak pivot would give a huge, indefinite
        // "inverse", so the block falls back to 1/diag and is counted (icqb_cg_last_blocks)
        if (nweak) s_bad = 1;
        __syncthreads();
        if (s_bad) {
            const double* s_diag = s_vec + 4 * kT;
            for (int e = threadIdx.x; e < kT * kT; e += kInvThreads) {
                const int i = e >> 7, j = e & (kT - 1);
                g[(long long)i * D.m + j] = (i == j) ? 1.0 / s_diag[i] : 0.0;
            }
            if (threadIdx.x == 0 && bad) atomicAdd(bad, 1);
            return;
        }
    } else if (nweak && weak) {
        atomicAdd(weak, nweak);   // inside a larger block weak pivots are reported only
    }
    regs_to_tile_symmetric(a, s_t, g, D.m);
}

// slab kc .. kc + 15 of a k-major operand X[k][c], c < WIDTH: into registers (WIDTH / 16 double2
// per thread), then into shared memory
template <int WIDTH>
struct SwSlab {
    static constexpr int kPairs = WIDTH * kSwK / 2 / 128;   // double2 per thread: 8 (128 wide), 4 (64)
    double2 v[kPairs];
    // X[k][c] = src[k * ld + c]
    __device__ __forceinline__ void load_rows(const double* src, long long ld, int kc, int tid) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + 128 * q, kk = idx / (WIDTH / 2), c = (idx % (WIDTH / 2)) * 2;
            v[q] = *reinterpret_cast<const double2*>(src + (long long)(kc + kk) * ld + c);
        }
    }
    // X[k][c] = src[c * ld + k]  (the transposed tile)
    __device__ __forceinline__ void load_cols(const double* src, long long ld, int kc, int tid) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + 128 * q, c = idx >> 3, kk = (idx & 7) * 2;
            v[q] = *reinterpret_cast<const double2*>(src + (long long)c * ld + kc + kk);
        }
    }
    template <int LD>
    __device__ __forceinline__ void store_rows(double (*dst)[LD], int tid, double sign) const {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + 128 * q, kk = idx / (WIDTH / 2), c = (idx % (WIDTH / 2)) * 2;
            *reinterpret_cast<double2*>(&dst[kk][c]) = make_double2(sign * v[q].x, sign * v[q].y);
        }
    }
    template <int LD>
    __device__ __forceinline__ void store_cols(double (*dst)[LD], int tid) const {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int idx = tid + 128 * q, c = idx >> 3, kk = (idx & 7) * 2;
            dst[kk][c] = v[q].x;
            dst[kk + 1][c] = v[q].y;
        }
    }
};

// one slab of 16: acc[i][j] += As[kk][rows of the thread] * Bs[kk][columns of the thread]
__device__ __forceinline__ void sw_multiply(const double (*As)[kSwM], const double (*Bs)[kSwLdB], int tx,
                                            int ty, double (&acc)[8][8]) {
#pragma unroll
    for (int kk = 0; kk < kSwK; ++kk) {
        double a[8], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(&As[kk][2 * tx + 32 * i]);
            a[2 * i] = v.x;
            a[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double2 v = *reinterpret_cast<const double2*>(&Bs[kk][8 * ty + 2 * i]);
            b[2 * i] = v.x;
            b[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// (b) W[:, c0 .. c0 + 63] = P V[:, c0 .. c0 + 63]; grid (m / 64, 1, blocks).  P is symmetric, so
// its k-major form is P itself.  The V chunk passes through shared memory anyway and is written
// to the V panel from there (k-major, whichever triangle it came from).
__global__ void __launch_bounds__(128, 2) k_sweep_panel(const DenseDesc* desc, double* dense, double* panels,
                                                        long long K) {
    __shared__ __align__(16) double As[2][kSwK][kSwM];
    __shared__ __align__(16) double Bs[2][kSwK][kSwLdB];
    const DenseDesc D = desc[blockIdx.z];
    if (K >= D.tiles) return;
    const long long m = D.m, c0 = (long long)blockIdx.x * kSwN, J = c0 / kT;
    if (J >= D.tiles || J == K) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double* base = dense + D.off;
    const double* P = base + (K * kT) * m + K * kT;
    // V[k][c0 + c]: row K of the block (J < K), or column K transposed (J > K)
    const bool fromRow = J < K;
    const double* Vsrc = fromRow ? base + (K * kT) * m + c0 : base + c0 * m + K * kT;
    double* Vp = panels + 2 * D.row0pad * kT;           // this block's V panel [128][m]
    double* Wp = Vp + kT * m;                           // and W panel

    SwSlab<kSwM> sa;
    SwSlab<kSwN> sb;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    sa.load_rows(P, m, 0, tid);
    if (fromRow) sb.load_rows(Vsrc, m, 0, tid); else sb.load_cols(Vsrc, m, 0, tid);
    sa.store_rows(As[0], tid, 1.0);
    if (fromRow) sb.store_rows(Bs[0], tid, 1.0); else sb.store_cols(Bs[0], tid);
    __syncthreads();
    for (int s = 0; s < kT / kSwK; ++s) {
        const int cur = s & 1;
        if (s + 1 < kT / kSwK) {
            sa.load_rows(P, m, (s + 1) * kSwK, tid);
            if (fromRow) sb.load_rows(Vsrc, m, (s + 1) * kSwK, tid); else sb.load_cols(Vsrc, m, (s + 1) * kSwK, tid);
        }
        // the V slab to its panel, k-major: 16 x 64 doubles, four double2 per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + 128 * q, kk = idx >> 5, c = (idx & 31) * 2;
            *reinterpret_cast<double2*>(Vp + (long long)(s * kSwK + kk) * m + c0 + c) =
                *reinterpret_cast<const double2*>(&Bs[cur][kk][c]);
        }
        sw_multiply(As[cur], Bs[cur], tx, ty, acc);
        if (s + 1 < kT / kSwK) {
            sa.store_rows(As[cur ^ 1], tid, 1.0);
            if (fromRow) sb.store_rows(Bs[cur ^ 1], tid, 1.0); else sb.store_cols(Bs[cur ^ 1], tid);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 2 * tx + 32 * (i >> 1) + (i & 1);
        double2* dst = reinterpret_cast<double2*>(Wp + (long long)row * m + c0 + 8 * ty);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_double2(acc[i][2 * j], acc[i][2 * j + 1]);
    }
}

// (c) A_IJ[:, n0 .. n0 + 63] += s_I V_I^T W_J; grid (T (T - 1) / 2 pairs x 2 halves, 1, blocks)
__global__ void __launch_bounds__(128, 2) k_sweep_update(const DenseDesc* desc, double* dense,
                                                         const double* panels, long long K) {
    __shared__ __align__(16) double As[2][kSwK][kSwM];
    __shared__ __align__(16) double Bs[2][kSwK][kSwLdB];
    const DenseDesc D = desc[blockIdx.z];
    if (K >= D.tiles) return;
    // pair (a >= b) over the tile indices without K
    const long long t = blockIdx.x >> 1;
    long long a = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (a * (a + 1) / 2 > t) --a;
    while ((a + 1) * (a + 2) / 2 <= t) ++a;
    const long long b = t - a * (a + 1) / 2;
    if (a >= D.tiles - 1) return;
    const long long I = a + (a >= K ? 1 : 0), J = b + (b >= K ? 1 : 0);
    const long long m = D.m, n0 = (long long)(blockIdx.x & 1) * kSwN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double* Vp = panels + 2 * D.row0pad * kT;
    const double* Wp = Vp + kT * m;
    const double* Asrc = Vp + I * kT;
    const double* Bsrc = Wp + J * kT + n0;
    const double sign = (I < K) ? 1.0 : -1.0;
    double* C = dense + D.off + (I * kT) * m + J * kT + n0;

    SwSlab<kSwM> sa;
    SwSlab<kSwN> sb;
    sa.load_rows(Asrc, m, 0, tid);
    sb.load_rows(Bsrc, m, 0, tid);
    // the accumulators start from C: 32 independent loads per thread, in flight with the first slab
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 2 * tx + 32 * (i >> 1) + (i & 1);
        const double2* src = reinterpret_cast<const double2*>(C + (long long)row * m + 8 * ty);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = src[j];
            acc[i][2 * j] = v.x;
            acc[i][2 * j + 1] = v.y;
        }
    }
    sa.store_rows(As[0], tid, sign);
    sb.store_rows(Bs[0], tid, 1.0);
    __syncthreads();
    for (int s = 0; s < kT / kSwK; ++s) {
        const int cur = s & 1;
        if (s + 1 < kT / kSwK) {
            sa.load_rows(Asrc, m, (s + 1) * kSwK, tid);
            sb.load_rows(Bsrc, m, (s + 1) * kSwK, tid);
        }
        sw_multiply(As[cur], Bs[cur], tx, ty, acc);
        if (s + 1 < kT / kSwK) {
            sa.store_rows(As[cur ^ 1], tid, sign);
            sb.store_rows(Bs[cur ^ 1], tid, 1.0);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 2 * tx + 32 * (i >> 1) + (i & 1);
        double2* dst = reinterpret_cast<double2*>(C + (long long)row * m + 8 * ty);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_double2(acc[i][2 * j], acc[i][2 * j + 1]);
    }
}

// (d) row K: A_KJ = W_J (J < K); column K: A_IK = -W_I^T (I > K).  Grid (m / 128, 1, blocks),
// 256 threads, 32 x 32 sub-tiles through shared memory for the transposed copies.
__global__ void __launch_bounds__(256) k_sweep_finish(const DenseDesc* desc, double* dense,
                                                      const double* panels, long long K) {
    __shared__ double s_t[32][33];
    const DenseDesc D = desc[blockIdx.z];
    const long long J = blockIdx.x;
    if (K >= D.tiles || J >= D.tiles || J == K) return;
    const long long m = D.m;
    const double* W = panels + 2 * D.row0pad * kT + kT * m + J * kT;   // W_J[k][j], pitch m
    double* base = dense + D.off;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 32 x 8
    if (J < K) {
        double* dst = base + (K * kT) * m + J * kT;
        for (int e = threadIdx.x; e < kT * kT; e += 256) {
            const int k = e >> 7, j = e & (kT - 1);
            dst[(long long)k * m + j] = W[(long long)k * m + j];
        }
        return;
    }
    double* dst = base + (J * kT) * m + K * kT;   // A_JK[j][k] = -W_J[k][j]
    for (int sub = 0; sub < 16; ++sub) {
        const int k0 = (sub >> 2) * 32, j0 = (sub & 3) * 32;
        for (int r = ly; r < 32; r += 8) s_t[r][lx] = W[(long long)(k0 + r) * m + j0 + lx];
        __syncthreads();
        for (int r = ly; r < 32; r += 8) dst[(long long)(j0 + r) * m + k0 + lx] = -s_t[lx][r];
        __syncthreads();
    }
}

// the upper triangle from the lower one; grid (mmax / 128, mmax, blocks)
__global__ void __launch_bounds__(kT) k_dense_mirror(const DenseDesc* desc, double* dense) {
    const DenseDesc D = desc[blockIdx.z];
    const long long i = blockIdx.y;
    const long long j = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= D.m || j >= D.m || j <= i) return;
    double* a = dense + D.off;
    a[i * D.m + j] = a[j * D.m + i];
}

// the block that holds compact row cr (row0 ascending): binary search -- a partition into single
// tiles has as many blocks as tile rows
__device__ __forceinline__ int find_block(const DenseDesc* desc, int nblocks, long long cr) {
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (desc[mid].row0 <= cr) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// w = C^-1 r on the rows of the dense blocks: HBM-bound (the inverse blocks are streamed once
// per iteration: 1024 * blockTiles * N bytes).  One warp per (valid) row; a lane reads 16-byte
// pairs of the row, 64 columns apart, eight of them in flight before the first is used (a
// single dependent 8-byte load per iteration reached 0.8 TB/s in round 2's first run), four
// partial sums per lane joined in a fixed order, then a butterfly.  The multiplied vector
// comes through L1 (every row of a block re-reads the same <= 64 KB).
// Grid ceil(rows / 8) CTAs of 256 threads.
__global__ void __launch_bounds__(256) k_dense_apply(const DenseDesc* desc, int nblocks,
                                                     long long totalRows, const double* dense,
                                                     const double* x, double* y, const int* skip) {
    if (skip && *skip) return;
    const int lane = threadIdx.x & 31;
    const long long cr = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (cr >= totalRows) return;   // whole warp
    const DenseDesc D = desc[find_block(desc, nblocks, cr)];
    const long long lr = cr - D.row0, g0 = D.t0 * kT;
    const double* row = dense + D.off + lr * D.m;   // 16-byte aligned: D.off and D.m are even
    const double* xv = x + g0;
    const long long valid = D.valid;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    long long c = 2 * lane;
    // full groups of eight pairs (512 columns per warp pass)
    for (; c + 7 * 64 + 1 < valid; c += 8 * 64) {
        double2 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            a[u] = __ldcs(reinterpret_cast<const double2*>(row + c + 64 * u));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc[u & 3] = fma(a[u].x, __ldg(xv + c + 64 * u), acc[u & 3]);
            acc[u & 3] = fma(a[u].y, __ldg(xv + c + 64 * u + 1), acc[u & 3]);
        }
    }
    for (; c < valid; c += 64) {
        const double2 a = __ldcs(reinterpret_cast<const double2*>(row + c));
        acc[0] = fma(a.x, __ldg(xv + c), acc[0]);
        if (c + 1 < valid) acc[1] = fma(a.y, __ldg(xv + c + 1), acc[1]);
    }
    double sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) y[g0 + lr] = sum;
}

// Sharded dense blocks (several ranks): a block whose tile rows all live on ONE rank is
// extracted, inverted and applied by that rank only; the pieces of w travel in one all-gather
// per iteration.  GatherDesc: rows [row, row + count) of w come from (or go to) position src of
// rank `owner`'s send buffer.
struct GatherDesc {
    long long row, count, src;
    int owner, pad;
};

// own pieces of w -> the send buffer
__global__ void __launch_bounds__(256) k_w_pack(const GatherDesc* g, int rank, const double* w,
                                                double* send, const int* skip) {
    if (skip && *skip) return;
    const GatherDesc D = g[blockIdx.y];
    if (D.owner != rank) return;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < D.count; i += (long long)gridDim.x * 256)
        send[D.src + i] = w[D.row + i];
}

// the other ranks' pieces, out of the gathered buffers [nranks][slot], into w
__global__ void __launch_bounds__(256) k_w_unpack(const GatherDesc* g, int rank, long long slot,
                                                  const double* recv, double* w, const int* skip) {
    if (skip && *skip) return;
    const GatherDesc D = g[blockIdx.y];
    if (D.owner == rank) return;
    const double* src = recv + (long long)D.owner * slot + D.src;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < D.count; i += (long long)gridDim.x * 256)
        w[D.row + i] = src[i];
}

// mixed-precision fields of the state (switch_thr 0: float32 products off)
__global__ void k_cg_mixed(CgState* s, double switch_thr) {
    s->lowp = 0;
    s->n32 = 0;
    s->switch_thr = switch_thr;
    s->skip64 = s->done;
    s->skip32 = 1;
}

// first half of k_cg_resid: the vector update only
__global__ void __launch_bounds__(kT) k_cg_vec(int mode, int k, int replace, const double* b,
                                               const double* scale, double* x, CgBuf B) {
    CgState* S = B.state;
    if (mode == kUpdate && S->done) return;
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= B.n) return;
    if (mode == kUpdate) {
        const double alpha = S->rho[k & 1] / *B.pz;
        B.r[i] = fma(-alpha, B.z[i], B.r[i]);
        x[i] = fma(alpha, B.p[i], x[i]);
    } else if (mode == kInit) {
        const double s = scale ? scale[i] : 1.0;
        B.s[i] = s;
        B.r[i] = s * b[i] - B.z[i];
        B.p[i] = 0.0;
    } else if (replace) {
        B.r[i] = B.s[i] * (b[i] - B.z[i]);
    }
}

// second half of k_cg_resid: w = M^-1 r (cap tiles: already written by the cap GEMV),
// partials, and the closing step by the last CTA -- the same arithmetic as k_cg_resid
__global__ void __launch_bounds__(kVThreads) k_cg_close(int mode, int k, int replace,
                                                        const double* b, CgBuf B,
                                                        const int* tileDense) {
    __shared__ double s_r[kT];
    __shared__ double s_acc[kGroups][kT];
    __shared__ double s_w4[5][4];
    __shared__ double s_fin[3 * kVThreads];
    __shared__ int s_last;
    CgState* S = B.state;
    if (mode == kUpdate && S->done) return;
    const int c = threadIdx.x, g = threadIdx.y, tid = g * kT + c;
    const long long T = blockIdx.x;
    const long long i = T * kT + c;
    const bool own = (g == 0) && (i < B.n);
    const bool capTile = tileDense[T] >= 0;    // uniform: the tile lies in a dense block
    const bool needW = (mode != kTrue) || replace;                           // uniform
    double r = 0.0, ru = 0.0, a1 = 0.0, a2 = 0.0, w = 0.0;
    double m[kGroupRows];
    if (needW && !capTile) {
        const double* M = B.Minv + T * kBlockElems + (long long)(g * kGroupRows) * kT + c;
#pragma unroll
        for (int rr = 0; rr < kGroupRows; ++rr) m[rr] = __ldg(M + rr * kT);
    }
    if (own) {
        if (mode == kTrue) {
            const double t = b[i] - B.z[i];
            a1 = t * t;
            a2 = fabs(t);
            if (replace) {
                r = B.r[i];
                ru = t;
            }
        } else {
            r = B.r[i];
            ru = r / B.s[i];
            if (mode == kInit) {
                const double bi = b[i];
                a1 = bi * bi;
                a2 = fabs(bi);
            }
        }
    }
    if (needW) {
        if (!capTile) {
            if (g == 0) s_r[c] = r;   // zero beyond n
            __syncthreads();
            double acc = 0.0;
#pragma unroll
            for (int rr = 0; rr < kGroupRows; ++rr) acc = fma(m[rr], s_r[g * kGroupRows + rr], acc);
            s_acc[g][c] = acc;
            __syncthreads();
            if (own) {
                w = ((s_acc[0][c] + s_acc[1][c]) + (s_acc[2][c] + s_acc[3][c])) +
                    ((s_acc[4][c] + s_acc[5][c]) + (s_acc[6][c] + s_acc[7][c]));
                B.w[i] = w;
            }
        } else if (own) {
            w = B.w[i];   // C^-1 r of this cap, written by the GEMV before this kernel
        }
    }
    if (g == 0) {
        warp_part<false>(r * w, s_w4[0], tid);
        warp_part<false>(ru * ru, s_w4[1], tid);
        warp_part<true>(fabs(ru), s_w4[2], tid);
        warp_part<false>(a1, s_w4[3], tid);
        warp_part<true>(a2, s_w4[4], tid);
    }
    __syncthreads();
    if (tid == 0) {
        if (needW) B.part_rw[T] = join4<false>(s_w4[0]);
        if (mode != kTrue) {
            B.part_res[T] = join4<false>(s_w4[1]);
            B.part_rmx[T] = join4<true>(s_w4[2]);
        }
        if (mode != kUpdate) {
            B.part_a[T] = join4<false>(s_w4[3]);
            B.part_b[T] = join4<true>(s_w4[4]);
        }
        s_last = take_ticket(&S->ticket_res, gridDim.x) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;

    __threadfence();
    const int nP = (int)gridDim.x;
    double rho = 0.0, res2 = 0.0, resmax = 0.0, fa = 0.0, fb = 0.0;
    if (mode != kTrue) {
        final_reduce3(B.part_rw, B.part_res, B.part_rmx, nP, s_fin, tid, kVThreads, &rho, &res2,
                      &resmax);
    } else if (replace) {
        rho = final_reduce<false>(B.part_rw, nP, s_fin, tid, kVThreads);
    }
    if (mode != kUpdate) {
        fa = final_reduce<false>(B.part_a, nP, s_fin, tid, kVThreads);
        fb = final_reduce<true>(B.part_b, nP, s_fin, tid, kVThreads);
    }
    if (tid == 0) {
        if (mode == kUpdate) {
            S->rho[(k + 1) & 1] = rho;
            S->beta = rho / S->rho[k & 1];
            S->res2 = res2;
            S->resmax = resmax;
            S->iters += 1;
            if ((S->use_max ? resmax : res2) <= S->thr) S->done = 1;
            // mixed precision: this iteration's product was a float32 one if lowp was set;
            // from the first residual below switch_thr on, the products read the float32 copy
            if (S->lowp) S->n32 += 1;
            if (S->switch_thr > 0.0 && resmax <= S->switch_thr) S->lowp = 1;
            S->skip64 = S->done | S->lowp;
            S->skip32 = S->done | (S->lowp ^ 1);
        } else if (mode == kInit) {
            S->rho[0] = rho;
            S->beta = 0.0;
            S->res2 = res2;
            S->resmax = resmax;
            S->bb = fa;
            S->bmax = fb;
        } else {
            S->tru2 = fa;
            S->trumax = fb;
            if (replace) {
                S->rho[k & 1] = rho;
                S->beta = (k > 0) ? rho / S->rho[(k - 1) & 1] : 0.0;
            }
        }
        S->ticket_res = 0;
        __threadfence();
    }
}

// One kernel instead of k_cg_vec + k_dense_apply + k_cg_close for an iteration of the common
// case (one rank, every tile in a dense block): CTA = kUpRows consecutive rows of one block.
//   * every CTA forms r_new = r_old - alpha z for the COLUMNS of its block in shared memory (the
//     multiplied vector of the block product; r is double-buffered by the parity of the
//     iteration, so no CTA reads entries another one has already replaced) and writes back x
//     and r_new for its own rows;
//   * each warp streams its rows of the inverse block against that vector, two rows at a time
//     (16 loads of 16 bytes in flight per lane), w = C^-1 r_new;
//   * partials of rho = r.w, |r/s|^2, max|r/s| per CTA; the CTA that finishes last adds them in
//     a fixed order and closes the iteration exactly as k_cg_close does.
// Columns of a block beyond n multiply zeros (the vector is zero-filled up to m; the inverse of
// the identity-padded block is finite there), so the inner loop has no bounds checks.
constexpr int kUpRows = 16;      // two rows per warp.  Measured at config C2 (B200, solve of 92 iterations): 32 rows
                                 // per CTA 33.5 ms, 16 rows 32.6 ms, 8 rows (one per warp, 2,485 CTAs that each
                                 // stage the vector) 34.2 ms
constexpr int kUpThreads = 256;
static_assert(kUpRows == 2 * (kUpThreads / 32), "every warp multiplies one pair of rows; warp 0 closes the CTA's rows");

// two rows of an inverse block against the staged vector, columns [c, m) of this lane's pairs:
// eight 16-byte loads per row in flight
__device__ __forceinline__ void up_row_pair(const double* __restrict__ rowA, const double* __restrict__ rowB,
                                            const double* s_r, long long m, long long c, double (&acc)[4]) {
    for (; c + 7 * 64 < m; c += 8 * 64) {
        double2 va[8], vb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            va[u] = __ldcs(reinterpret_cast<const double2*>(rowA + c + 64 * u));
            vb[u] = __ldcs(reinterpret_cast<const double2*>(rowB + c + 64 * u));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 x = *reinterpret_cast<const double2*>(s_r + c + 64 * u);
            acc[0] = fma(va[u].x, x.x, acc[0]);
            acc[1] = fma(va[u].y, x.y, acc[1]);
            acc[2] = fma(vb[u].x, x.x, acc[2]);
            acc[3] = fma(vb[u].y, x.y, acc[3]);
        }
    }
    for (; c < m; c += 64) {
        const double2 va = __ldcs(reinterpret_cast<const double2*>(rowA + c));
        const double2 vb = __ldcs(reinterpret_cast<const double2*>(rowB + c));
        const double2 x = *reinterpret_cast<const double2*>(s_r + c);
        acc[0] = fma(va.x, x.x, acc[0]);
        acc[1] = fma(va.y, x.y, acc[1]);
        acc[2] = fma(vb.x, x.x, acc[2]);
        acc[3] = fma(vb.y, x.y, acc[3]);
    }
}

__global__ void __launch_bounds__(kUpThreads) k_cg_update_dense(const DenseDesc* desc, int nblocks,
                                                                const double* dense, int k, double* x, CgBuf B,
                                                                const double* r_old, double* r_new,
                                                                double* part) {
    extern __shared__ __align__(16) double s_r[];   // [m of the largest block]
    __shared__ double s_w[kUpRows];
    __shared__ double s_fin[3 * kUpThreads];
    __shared__ int s_last;
    CgState* S = B.state;
    if (S->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long cr0 = (long long)blockIdx.x * kUpRows;
    const DenseDesc D = desc[find_block(desc, nblocks, cr0)];
    const long long lr0 = cr0 - D.row0, g0 = D.t0 * kT, m = D.m, valid = D.valid;
    // this warp's two rows; their first 512 columns are requested BEFORE the vector is staged, so
    // that the matrix loads are already in flight while the CTA forms r_new
    const double* inv = dense + D.off;
    const int lw = 2 * warp;
    const long long lrA = lr0 + lw, lrB = lrA + 1;
    const bool active = lrA < valid;                              // warp-uniform
    const long long rb = lrB < valid ? lrB : lrA;                 // (a ragged last row: row A twice)
    const double* rowA = inv + (active ? lrA : 0) * m;
    const double* rowB = inv + (active ? rb : 0) * m;
    const long long c0 = 2 * lane;
    const bool head = active && (c0 + 7 * 64 < m);               // uniform: m is a multiple of 128
    double2 va[8], vb[8];
    if (head) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            va[u] = __ldcs(reinterpret_cast<const double2*>(rowA + c0 + 64 * u));
            vb[u] = __ldcs(reinterpret_cast<const double2*>(rowB + c0 + 64 * u));
        }
    }
    const double alpha = S->rho[k & 1] / *B.pz;
    for (long long c = tid; c < m; c += kUpThreads)
        s_r[c] = (c < valid) ? fma(-alpha, B.z[g0 + c], r_old[g0 + c]) : 0.0;
    __syncthreads();
    if (tid < kUpRows && lr0 + tid < valid) {
        const long long i = g0 + lr0 + tid;
        x[i] = fma(alpha, B.p[i], x[i]);
        r_new[i] = s_r[lr0 + tid];
    }
    if (active) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        long long c = c0;
        if (head) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double2 xv = *reinterpret_cast<const double2*>(s_r + c0 + 64 * u);
                acc[0] = fma(va[u].x, xv.x, acc[0]);
                acc[1] = fma(va[u].y, xv.y, acc[1]);
                acc[2] = fma(vb[u].x, xv.x, acc[2]);
                acc[3] = fma(vb[u].y, xv.y, acc[3]);
            }
            c += 8 * 64;
        }
        up_row_pair(rowA, rowB, s_r, m, c, acc);
        double wa = acc[0] + acc[1], wb = acc[2] + acc[3];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            wa += __shfl_xor_sync(0xffffffffu, wa, off);
            wb += __shfl_xor_sync(0xffffffffu, wb, off);
        }
        if (lane == 0) {
            B.w[g0 + lrA] = wa;
            s_w[lw] = wa;
            if (lrB < valid) {
                B.w[g0 + lrB] = wb;
                s_w[lw + 1] = wb;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        double rw = 0.0, r2 = 0.0, rm = 0.0;
        const long long lr = lr0 + lane;
        if (lane < kUpRows && lr < valid) {
            const double r = s_r[lr];
            const double ru = r / B.s[g0 + lr];
            rw = r * s_w[lane];
            r2 = ru * ru;
            rm = fabs(ru);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            rw += __shfl_xor_sync(0xffffffffu, rw, off);
            r2 += __shfl_xor_sync(0xffffffffu, r2, off);
            rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, off));
        }
        if (lane == 0) {
            const long long nP = gridDim.x;
            part[blockIdx.x] = rw;
            part[nP + blockIdx.x] = r2;
            part[2 * nP + blockIdx.x] = rm;
            s_last = take_ticket(&S->ticket_res, gridDim.x) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nP = (int)gridDim.x;
    double rho = 0.0, res2 = 0.0, resmax = 0.0;
    final_reduce3(part, part + nP, part + 2 * nP, nP, s_fin, tid, kUpThreads, &rho, &res2, &resmax);
    if (tid == 0) {
        S->rho[(k + 1) & 1] = rho;
        S->beta = rho / S->rho[k & 1];
        S->res2 = res2;
        S->resmax = resmax;
        S->iters += 1;
        if ((S->use_max ? resmax : res2) <= S->thr) S->done = 1;
        if (S->lowp) S->n32 += 1;
        if (S->switch_thr > 0.0 && resmax <= S->switch_thr) S->lowp = 1;
        S->skip64 = S->done | S->lowp;
        S->skip32 = S->done | (S->lowp ^ 1);
        S->ticket_res = 0;
        __threadfence();
    }
}

// The same step for OVERLAPPING blocks (icqb_cg_set_block_extents; one rank):
//     w = sum_q R_q^T D_q A_q^-1 D_q R_q r_new,   D_q = 1 / sqrt(extents that hold the row).
// CTA = kUpRows consecutive rows of one EXTENT (the extents are laid one after the other, each
// padded to a multiple of kUpRows: `ctaBlock` gives the block of a CTA).  The staged vector is
// D r_new over the columns of the extent; a row's result is weighted and ADDED to w, which the
// host zeroes before the launch -- a row lies in two extents at most, so the sum of its two
// contributions does not depend on their order and the result is reproducible.  x and r are
// written, and the residual norms taken, by the CTAs of the block a row belongs to (its core
// rows); rho = r.w is summed extent by extent as (D r).(A_q^-1 D r).
// APPLY_ONLY: w = M^-1 r for the given r (set-up and residual replacement): no update, no
// partials, no closing step -- k_cg_close follows.
template <bool APPLY_ONLY>
__global__ void __launch_bounds__(kUpThreads) k_cg_update_overlap(const DenseDesc* desc, const int* ctaBlock,
                                                                  const double* dense, const double* dcover,
                                                                  int k, double* x, CgBuf B,
                                                                  const double* r_old, double* r_new,
                                                                  double* part) {
    extern __shared__ __align__(16) double s_r[];   // [m of the largest extent]
    __shared__ double s_w[kUpRows];
    __shared__ double s_fin[3 * kUpThreads];
    __shared__ int s_last;
    CgState* S = B.state;
    if (!APPLY_ONLY && S->done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DenseDesc D = desc[ctaBlock[blockIdx.x]];
    const long long lr0 = (long long)blockIdx.x * kUpRows - D.urow0;   // first row of the CTA in its extent
    const long long g0 = D.t0 * kT, m = D.m, valid = D.valid;
    const double* inv = dense + D.off;
    const int lw = 2 * warp;
    const long long lrA = lr0 + lw, lrB = lrA + 1;
    const bool active = lrA < valid;                              // warp-uniform
    const long long rb = lrB < valid ? lrB : lrA;
    const double* rowA = inv + (active ? lrA : 0) * m;
    const double* rowB = inv + (active ? rb : 0) * m;
    const long long c0 = 2 * lane;
    const bool head = active && (c0 + 7 * 64 < m);               // uniform: m is a multiple of 128
    double2 va[8], vb[8];
    if (head) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            va[u] = __ldcs(reinterpret_cast<const double2*>(rowA + c0 + 64 * u));
            vb[u] = __ldcs(reinterpret_cast<const double2*>(rowB + c0 + 64 * u));
        }
    }
    const double alpha = APPLY_ONLY ? 0.0 : S->rho[k & 1] / *B.pz;
    for (long long c = tid; c < m; c += kUpThreads) {
        double v = 0.0;
        if (c < valid) {
            const double rn = APPLY_ONLY ? r_old[g0 + c] : fma(-alpha, B.z[g0 + c], r_old[g0 + c]);
            v = dcover[g0 + c] * rn;
        }
        s_r[c] = v;
    }
    __syncthreads();
    // the rows of this CTA that belong to the block itself: x, r, residual norms
    double ruOwn = 0.0;
    bool own = false;
    if (!APPLY_ONLY && tid < kUpRows && lr0 + tid < valid) {
        const long long i = g0 + lr0 + tid;
        own = (i >= D.core0 && i < D.core1);
        if (own) {
            const double rn = fma(-alpha, B.z[i], r_old[i]);
            x[i] = fma(alpha, B.p[i], x[i]);
            r_new[i] = rn;
            ruOwn = rn / B.s[i];
        }
    }
    if (active) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        long long c = c0;
        if (head) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double2 xv = *reinterpret_cast<const double2*>(s_r + c0 + 64 * u);
                acc[0] = fma(va[u].x, xv.x, acc[0]);
                acc[1] = fma(va[u].y, xv.y, acc[1]);
                acc[2] = fma(vb[u].x, xv.x, acc[2]);
                acc[3] = fma(vb[u].y, xv.y, acc[3]);
            }
            c += 8 * 64;
        }
        up_row_pair(rowA, rowB, s_r, m, c, acc);
        double ya = acc[0] + acc[1], yb = acc[2] + acc[3];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            ya += __shfl_xor_sync(0xffffffffu, ya, off);
            yb += __shfl_xor_sync(0xffffffffu, yb, off);
        }
        if (lane == 0) {
            atomicAdd(&B.w[g0 + lrA], dcover[g0 + lrA] * ya);
            s_w[lw] = ya;
            if (lrB < valid) {
                atomicAdd(&B.w[g0 + lrB], dcover[g0 + lrB] * yb);
                s_w[lw + 1] = yb;
            }
        }
    }
    if (APPLY_ONLY) return;
    __syncthreads();
    if (warp == 0) {
        double rw = 0.0, r2 = 0.0, rm = 0.0;
        const long long lr = lr0 + lane;
        if (lane < kUpRows && lr < valid) rw = s_r[lr] * s_w[lane];   // (D r) . (A_q^-1 D r)
        if (own) {
            r2 = ruOwn * ruOwn;
            rm = fabs(ruOwn);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            rw += __shfl_xor_sync(0xffffffffu, rw, off);
            r2 += __shfl_xor_sync(0xffffffffu, r2, off);
            rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, off));
        }
        if (lane == 0) {
            const long long nP = gridDim.x;
            part[blockIdx.x] = rw;
            part[nP + blockIdx.x] = r2;
            part[2 * nP + blockIdx.x] = rm;
            s_last = take_ticket(&S->ticket_res, gridDim.x) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nP = (int)gridDim.x;
    double rho = 0.0, res2 = 0.0, resmax = 0.0;
    final_reduce3(part, part + nP, part + 2 * nP, nP, s_fin, tid, kUpThreads, &rho, &res2, &resmax);
    if (tid == 0) {
        S->rho[(k + 1) & 1] = rho;
        S->beta = rho / S->rho[k & 1];
        S->res2 = res2;
        S->resmax = resmax;
        S->iters += 1;
        if ((S->use_max ? resmax : res2) <= S->thr) S->done = 1;
        if (S->lowp) S->n32 += 1;
        if (S->switch_thr > 0.0 && resmax <= S->switch_thr) S->lowp = 1;
        S->skip64 = S->done | S->lowp;
        S->skip32 = S->done | (S->lowp ^ 1);
        S->ticket_res = 0;
        __threadfence();
    }
}

}  // namespace

int cg_solve_caps(icqb_ctx* ctx, uint64_t dA_, int64_t n, int64_t ld, int storage,
                  uint64_t drowscale_, uint64_t db_, uint64_t dx_, double tol, int rel,
                  int64_t maxit, int64_t* iters, double* resid2, double* residinf) {
    // (arguments were validated by icqb_cg_solve_dev)
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double* A = reinterpret_cast<const double*>(dA_);
    const double* scale = reinterpret_cast<const double*>(drowscale_);
    const double* b = reinterpret_cast<const double*>(db_);
    double* x = reinterpret_cast<double*>(dx_);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long chunk = (n + P - 1) / P;
    if (maxit < 0) maxit = n;
    cudaStream_t st = ctx->stream;
    const long long nT = (n + kT - 1) / kT;
    if (storage == 1 &&
        ((ctx->r1 % kT != 0 && ctx->r1 != n) || (ctx->q1 > ctx->q0 && ctx->q1 % kT != 0 && ctx->q1 != n)))
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: the block preconditioner needs row blocks that end on "
                         "multiples of 128 (or at n)");

    // the partition of the tile rows into diagonal blocks: given explicitly
    // (icqb_cg_set_block_partition), or two caps around single tiles (icqb_cg_set_cap_tiles)
    std::vector<long long> starts;
    if (!ctx->block_starts.empty()) {
        for (long long t : ctx->block_starts)
            if (t < nT) starts.push_back(t);
    } else {
        const long long head = std::min<long long>(std::max<long long>(ctx->cap_head_tiles, 0), nT);
        const long long tail = std::min<long long>(std::max<long long>(ctx->cap_tail_tiles, 0), nT - head);
        starts.push_back(0);
        for (long long t = std::max<long long>(head, 1); t <= nT - tail && t < nT; ++t) starts.push_back(t);
    }
    if (starts.empty() || starts[0] != 0) starts.insert(starts.begin(), 0);
    // every dense block (more than one tile), in index order
    struct BlockRange { long long t0, t1; };
    std::vector<BlockRange> ranges;
    std::vector<int> htile((size_t)nT, -1);
    for (size_t q = 0; q < starts.size(); ++q) {
        const long long t0 = starts[q], t1 = (q + 1 < starts.size()) ? starts[q + 1] : nT;
        // (single tiles are dense blocks of one tile: same extraction, inversion and product)
        for (long long t = t0; t < t1; ++t) htile[(size_t)t] = (int)ranges.size();
        ranges.push_back({t0, t1});
    }
    const int nAll = (int)ranges.size();
    // overlapping extents (icqb_cg_set_block_extents): one rank only -- with several ranks every
    // block is applied by the rank that holds its rows, without overlap
    bool overlap = (P == 1) && ctx->block_ext0.size() == starts.size() && !ctx->block_starts.empty() &&
                   ctx->block_starts.size() == starts.size();
    std::vector<BlockRange> extents = ranges;
    if (overlap) {
        bool any = false;
        for (int bq = 0; bq < nAll; ++bq) {
            extents[(size_t)bq].t0 = std::min<long long>(ctx->block_ext0[(size_t)bq], ranges[(size_t)bq].t0);
            extents[(size_t)bq].t1 = std::min<long long>(std::max<long long>(ctx->block_ext1[(size_t)bq],
                                                                               ranges[(size_t)bq].t1), nT);
            any = any || extents[(size_t)bq].t0 != ranges[(size_t)bq].t0 || extents[(size_t)bq].t1 != ranges[(size_t)bq].t1;
        }
        overlap = any;
    }
    // Who holds a block?  A block whose tile rows all lie in this rank's rows is OWNED: it is
    // extracted, inverted and applied here only, and its piece of w is all-gathered.  Blocks
    // that straddle a boundary between ranks (none when the host layer aligns the partition with
    // the row blocks, icqRowPartition.sliverBlockPartition(cuts=...)) stay REPLICATED: extracted
    // by all-reduce, inverted and applied by every rank.  One small all-reduce tells every rank
    // the owner of every block.
    LowerLayout L{0, 0, 0, 0, 0, 0, 0};
    if (storage == 1) {
        int rcL = lower_layout(ctx, &L);
        if (rcL != ICQB_OK) return rcL;
    }
    const long long fr0 = std::min<long long>(n, (long long)rank * chunk);
    const long long fr1 = std::min<long long>(n, fr0 + chunk);
    auto holdsTileRow = [&](long long t) -> bool {
        if (storage == 1) return L.stripOf(t * kT) >= 0;
        return t * kT >= fr0 && std::min<long long>(n, (t + 1) * kT) <= fr1;
    };
    std::vector<int> owner((size_t)nAll, P == 1 ? 0 : -1);
    if (P > 1 && nAll > 0) {
        std::vector<double> vote(2 * (size_t)nAll, 0.0);
        for (int bq = 0; bq < nAll; ++bq) {
            bool all = true;
            for (long long t = ranges[bq].t0; t < ranges[bq].t1 && all; ++t) all = holdsTileRow(t);
            if (all) {
                vote[2 * bq] = 1.0;
                vote[2 * bq + 1] = rank + 1.0;
            }
        }
        int rcV = ensure_work(ctx, 2 * (long long)nAll + 64);
        if (rcV != ICQB_OK) return rcV;
        ICQB_CUDA(ctx, cudaMemcpyAsync(ctx->d_work, vote.data(), sizeof(double) * vote.size(),
                                       cudaMemcpyHostToDevice, st));
        rcV = comm_allreduce_sum(ctx, ctx->d_work, 2 * (long long)nAll, st);
        if (rcV != ICQB_OK) return rcV;
        ICQB_CUDA(ctx, cudaMemcpyAsync(vote.data(), ctx->d_work, sizeof(double) * vote.size(),
                                       cudaMemcpyDeviceToHost, st));
        ICQB_CUDA(ctx, cudaStreamSynchronize(st));
        for (int bq = 0; bq < nAll; ++bq)
            if (vote[2 * bq] == 1.0) owner[bq] = (int)llround(vote[2 * bq + 1]) - 1;
    }
    // local list: the blocks this rank inverts and applies (owned first, then replicated)
    std::vector<DenseDesc> hdesc;
    std::vector<GatherDesc> hgather;
    std::vector<long long> sendLen((size_t)P, 0);
    long long denseDoubles = 0, denseRows = 0, denseRowsPad = 0, kMaxTiles = 0, repDoubles = 0, repOff = 0;
    long long updRowsPad = 0;
    int nOwn = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int bq = 0; bq < nAll; ++bq) {
            const bool mine = (owner[bq] == rank), rep = (owner[bq] < 0);
            if (pass == 0 && bq < nAll && owner[bq] >= 0) {
                // gather table: every owned block, whoever owns it
                GatherDesc G;
                G.row = ranges[bq].t0 * kT;
                G.count = std::min<long long>(n, ranges[bq].t1 * kT) - G.row;
                G.owner = owner[bq];
                G.pad = 0;
                G.src = sendLen[(size_t)owner[bq]];
                sendLen[(size_t)owner[bq]] += G.count;
                hgather.push_back(G);
            }
            if ((pass == 0 && !mine) || (pass == 1 && !rep)) continue;
            DenseDesc D;
            D.t0 = extents[bq].t0;
            D.tiles = extents[bq].t1 - extents[bq].t0;
            D.m = D.tiles * kT;
            D.valid = std::min<long long>(n, extents[bq].t1 * kT) - D.t0 * kT;
            D.off = denseDoubles;
            D.row0 = denseRows;
            D.row0pad = denseRowsPad;
            D.core0 = ranges[bq].t0 * kT;
            D.core1 = std::min<long long>(n, ranges[bq].t1 * kT);
            D.urow0 = updRowsPad;
            updRowsPad += (D.valid + kUpRows - 1) / kUpRows * kUpRows;
            denseDoubles += D.m * D.m;
            denseRows += D.valid;
            denseRowsPad += D.m;
            kMaxTiles = std::max(kMaxTiles, D.tiles);
            hdesc.push_back(D);
            if (pass == 0) ++nOwn;
            else repDoubles += D.m * D.m;
        }
        if (pass == 0) repOff = denseDoubles;
    }
    const bool sharded = (P > 1) && !hgather.empty();
    long long slot = 0;
    for (long long v : sendLen) slot = std::max(slot, v);
    const long long gatherDoubles = sharded ? (long long)(P + 1) * slot : 0;
    const long long gdescDoubles = ((long long)sizeof(GatherDesc) * std::max<size_t>(hgather.size(), 1) + 7) / 8;
    const int nDense = (int)hdesc.size();
    if (kMaxTiles > 64) return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: dense blocks of at most 64 tiles");
    const long long descDoubles = ((long long)sizeof(DenseDesc) * std::max(nDense, 1) + 7) / 8;
    const long long tileDoubles = (4 * nT + 7) / 8;

    // work space
    const long long nfull = chunk * P;
    const long long stateDoubles = (sizeof(CgState) + 7) / 8;
    const long long small = 8 * n + 8 + nfull + 6 * nT + stateDoubles;
    const long long panelDoubles = 2 * kT * denseRowsPad;   // V and W panels of the block sweep
    // The fused update kernel (k_cg_update_dense) serves the common case: one rank, every tile in
    // a dense block, and only the last block ragged (its CTAs never straddle two blocks).
    bool fusedUpdate = (P == 1) && nDense > 0;
    for (long long t = 0; t < nT && fusedUpdate; ++t) fusedUpdate = htile[(size_t)t] >= 0;
    for (int q = 0; q + 1 < nDense && fusedUpdate; ++q) fusedUpdate = hdesc[(size_t)q].valid == hdesc[(size_t)q].m;
    {
        static const bool off = [] {
            const char* e = getenv("ICQB_CG_UNFUSED");   // debugging aid: the three-kernel update
            return e && e[0] == '1';
        }();
        if (off) fusedUpdate = false;
    }
    if (overlap) fusedUpdate = true;   // (its own kernel; shares the double-buffered residual)
    const long long nUpd = overlap ? updRowsPad / kUpRows : (denseRows + kUpRows - 1) / kUpRows;
    // partials of the update kernel; overlapping form: + CTA -> block table, + 1/sqrt(cover) per row
    const long long updDoubles = fusedUpdate ? 3 * nUpd + (overlap ? (nUpd + 1) / 2 + n + 2 : 0) : 0;
    const long long need = small + denseDoubles + panelDoubles + descDoubles + tileDoubles +
                           gatherDoubles + gdescDoubles + updDoubles + 68;
    int rc = ensure_work(ctx, need);
    if (rc != ICQB_OK) return rc;
    CgBuf B;
    double* wp = ctx->d_work;
    B.p = wp; wp += n;
    B.r = wp; wp += n;
    B.w = wp; wp += n;
    B.z = wp; wp += n + 8;
    B.minv = wp; wp += n;
    B.s = wp; wp += n;
    double* r_alt = wp; wp += n;   // second residual buffer of the fused update (cg.cu keeps diag(A) here)
    double* inv_area = wp; wp += n;
    double* gathered = wp; wp += nfull;
    B.part_rw = wp; wp += nT;
    B.part_res = wp; wp += nT;
    B.part_rmx = wp; wp += nT;
    B.part_a = wp; wp += nT;
    B.part_b = wp; wp += nT;
    B.part_pz = wp; wp += nT;
    B.state = reinterpret_cast<CgState*>(wp); wp += stateDoubles;
    double* blocks = nullptr;   // (the 128 x 128 inverses of cg.cu: not used by this solver)
    if ((reinterpret_cast<uintptr_t>(wp) & 15) != 0) wp += 1;   // k_dense_apply reads 16-byte pairs
    double* dense = wp; wp += denseDoubles;
    double* panels = wp; wp += panelDoubles;
    DenseDesc* ddesc = reinterpret_cast<DenseDesc*>(wp); wp += descDoubles;
    int* dtile = reinterpret_cast<int*>(wp); wp += tileDoubles;
    double* wsend = wp; wp += sharded ? slot : 0;
    double* wrecv = wp; wp += sharded ? (long long)P * slot : 0;
    GatherDesc* dgather = reinterpret_cast<GatherDesc*>(wp); wp += gdescDoubles;
    double* updPart = wp; wp += updDoubles;
    int* updBlock = reinterpret_cast<int*>(updPart + 3 * nUpd);
    double* dcover = updPart + 3 * nUpd + (nUpd + 1) / 2 + 1;
    std::vector<int> hUpdBlock;
    std::vector<double> hCover;
    if (overlap) {
        hUpdBlock.resize((size_t)nUpd);
        hCover.assign((size_t)n, 0.0);
        for (int bq = 0; bq < nDense; ++bq) {
            const DenseDesc& D = hdesc[(size_t)bq];
            const long long q0 = D.urow0 / kUpRows, q1 = q0 + (D.valid + kUpRows - 1) / kUpRows;
            for (long long q = q0; q < q1; ++q) hUpdBlock[(size_t)q] = bq;
            for (long long i = D.t0 * kT; i < D.t0 * kT + D.valid; ++i) hCover[(size_t)i] += 1.0;
        }
        for (long long i = 0; i < n; ++i) hCover[(size_t)i] = 1.0 / sqrt(hCover[(size_t)i]);
    }
    double* rbuf[2] = {B.r, fusedUpdate ? r_alt : B.r};
    // the residual of iteration k lives in rbuf[k & 1] (fused update: double-buffered)
    auto atIter = [&](long long kq) {
        CgBuf c = B;
        c.r = rbuf[kq & 1];
        return c;
    };
    size_t updSmem = 0;
    if (fusedUpdate) {
        updSmem = sizeof(double) * (size_t)(kMaxTiles * kT);
        ICQB_CUDA(ctx, cudaFuncSetAttribute(k_cg_update_dense, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)updSmem));
        ICQB_CUDA(ctx, cudaFuncSetAttribute(k_cg_update_overlap<false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)updSmem));
        ICQB_CUDA(ctx, cudaFuncSetAttribute(k_cg_update_overlap<true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)updSmem));
    }
    B.Minv = blocks;
    B.n = n;

    MatVec mv;
    mv.ctx = ctx;
    mv.A = A;
    mv.n = n;
    mv.ld = ld;
    mv.storage = storage;
    mv.chunk = chunk;
    mv.r0 = std::min<long long>(n, (long long)rank * chunk);
    mv.r1 = std::min<long long>(n, mv.r0 + chunk);
    mv.area = ctx->d_area2;
    mv.inv_area = inv_area;
    mv.gathered = gathered;

    SymvFuse fuse;
    const bool fuseDir = (storage == 1);
    // peer exchange instead of the NCCL all-reduce of the iteration's product
    SymvPeer probe;
    const bool usePeer = fuseDir && P > 1 && comm_peer_next(ctx, n, &probe);
    B.pz = (fuseDir && P > 1 && !usePeer) ? B.z + n : &B.state->pz;
    if (fuseDir) {
        fuse.w = B.w;
        fuse.beta = &B.state->beta;
        fuse.p = B.p;
        fuse.pz_part = B.part_pz;
        fuse.pz = const_cast<double*>(B.pz);
        fuse.ticket = &B.state->ticket_dot;
    }

    cudaEvent_t evTotal0 = ctx->ev0, evTotal1 = ctx->ev1;
    cudaEvent_t evg0[2] = {nullptr, nullptr}, evg1[2] = {nullptr, nullptr}, evb[2] = {nullptr, nullptr};
    {
        // (created together: a failure leaves nothing behind)
        cudaError_t ee = cudaSuccess;
        for (int q = 0; q < 2 && ee == cudaSuccess; ++q) {
            ee = cudaEventCreate(&evg0[q]);
            if (ee == cudaSuccess) ee = cudaEventCreate(&evg1[q]);
            if (ee == cudaSuccess) ee = cudaEventCreateWithFlags(&evb[q], cudaEventDisableTiming);
        }
        if (ee != cudaSuccess) {
            for (int q = 0; q < 2; ++q) {
                if (evg0[q]) cudaEventDestroy(evg0[q]);
                if (evg1[q]) cudaEventDestroy(evg1[q]);
                if (evb[q]) cudaEventDestroy(evb[q]);
            }
            return check_cuda(ctx, ee, "cudaEventCreate");
        }
    }
    double gemv_ms = 0.0;
    long long gemv_count = 0;
    int status = ICQB_OK;
    const dim3 vblock(kT, kGroups);
    const unsigned vgrid = (unsigned)nT;
    const int* skip = reinterpret_cast<const int*>(B.state);   // &state->done
    CgState* hslot = reinterpret_cast<CgState*>(ctx->h_pinned);

    const bool mixedOn = (ctx->mixed_switch > 0.0) && storage == 1;   // float32 late products
    long long products32 = 0;

    // w = M^-1 r and the closing step, after r has been updated by k_cg_vec
    auto precondition = [&](int mode, int kk, int replace, const int* skipFlag) -> int {
        const CgBuf B = atIter(kk);   // (shadows the solver's: the residual buffer of iteration kk)
        const bool needW = (mode != kTrue) || replace;
        if (needW && overlap) {
            // overlapping blocks: the weighted contributions of the extents are added into a zeroed w
            int rco = check_cuda(ctx, cudaMemsetAsync(B.w, 0, sizeof(double) * (size_t)n, st), "cudaMemsetAsync w");
            if (rco != ICQB_OK) return rco;
            k_cg_update_overlap<true><<<(unsigned)nUpd, kUpThreads, updSmem, st>>>(
                ddesc, updBlock, dense, dcover, kk, nullptr, B, B.r, nullptr, nullptr);
            ctx->launches += 1;
        } else if (needW && denseRows > 0) {
            k_dense_apply<<<(unsigned)((denseRows + 7) / 8), 256, 0, st>>>(ddesc, nDense, denseRows, dense,
                                                                         B.r, B.w, skipFlag);
            ctx->launches += 1;
        }
        if (needW && sharded) {
            // the owners' pieces of w to every rank: pack, all-gather, unpack
            const dim3 ggrid(4, (unsigned)hgather.size());
            k_w_pack<<<ggrid, 256, 0, st>>>(dgather, rank, B.w, wsend, skipFlag);
            int rcg = comm_allgather(ctx, wsend, wrecv, slot, st);
            if (rcg != ICQB_OK) return rcg;
            k_w_unpack<<<ggrid, 256, 0, st>>>(dgather, rank, slot, wrecv, B.w, skipFlag);
            ctx->launches += 2;
        }
        k_cg_close<<<vgrid, vblock, 0, st>>>(mode, kk, replace, b, B, dtile);
        ctx->launches += 1;
        return check_cuda(ctx, cudaGetLastError(), "k_cg_close launch");
    };

    auto enqueue_batch = [&](long long kstart, long long nb, int slot) -> int {
        for (long long i = 0; i < nb; ++i) {
            const int kk = (int)((kstart + i) & 0x3fffffff);
            if (!fuseDir) {
                k_cg_dir<<<128, 256, 0, st>>>(B);
                ctx->launches += 1;
            }
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg0[slot], st));
            int r2;
            if (usePeer) {
                SymvPeer X;
                if (!comm_peer_next(ctx, n, &X)) return set_error(ctx, ICQB_ESTATE, "peer exchange lost");
                // the reduce kernel leaves this rank's p.z in a scratch slot; the sum over the
                // ranks is formed by the gather kernel
                SymvFuse f2 = fuse;
                f2.pz = B.z + n;
                if (mixedOn)
                    r2 = launch_symv_mixed(ctx, A, ctx->d_A32, B.p, mv.area, scale ? mv.area : nullptr,
                                           scale ? nullptr : mv.inv_area, B.z, st, skip,
                                           &B.state->skip64, &B.state->skip32, &f2, &X);
                else
                    r2 = launch_symv(ctx, A, B.p, mv.area, scale ? mv.area : nullptr,
                                     scale ? nullptr : mv.inv_area, B.z, st, skip, &f2, &X);
                if (r2 != ICQB_OK) return r2;
                ++mv.products;
                k_peer_gather<<<vgrid, kT, 0, st>>>(X, n, B.z, &B.state->pz, skip);
                ctx->launches += 1;
            } else if (mixedOn) {
                r2 = launch_symv_mixed(ctx, A, ctx->d_A32, B.p, mv.area, scale ? mv.area : nullptr,
                                       scale ? nullptr : mv.inv_area, B.z, st, skip, &B.state->skip64,
                                       &B.state->skip32, &fuse, nullptr);
                if (r2 != ICQB_OK) return r2;
                ++mv.products;
                r2 = comm_allreduce_sum(ctx, B.z, n + 1, st);
                if (r2 != ICQB_OK) return r2;
            } else {
                r2 = mv.apply(B.p, B.z, scale, st, skip, fuseDir ? &fuse : nullptr, fuseDir ? 1 : 0);
                if (r2 != ICQB_OK) return r2;
            }
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg1[slot], st));
            if (!fuseDir) {
                k_cg_dot<<<vgrid, kT, 0, st>>>(B);
                ctx->launches += 1;
            }
            if (overlap) {
                r2 = check_cuda(ctx, cudaMemsetAsync(B.w, 0, sizeof(double) * (size_t)n, st), "cudaMemsetAsync w");
                if (r2 != ICQB_OK) return r2;
                k_cg_update_overlap<false><<<(unsigned)nUpd, kUpThreads, updSmem, st>>>(
                    ddesc, updBlock, dense, dcover, kk, x, B, rbuf[kk & 1], rbuf[(kk + 1) & 1], updPart);
                ctx->launches += 1;
                r2 = check_cuda(ctx, cudaGetLastError(), "k_cg_update_overlap launch");
            } else if (fusedUpdate) {
                k_cg_update_dense<<<(unsigned)nUpd, kUpThreads, updSmem, st>>>(
                    ddesc, nDense, dense, kk, x, B, rbuf[kk & 1], rbuf[(kk + 1) & 1], updPart);
                ctx->launches += 1;
                r2 = check_cuda(ctx, cudaGetLastError(), "k_cg_update_dense launch");
            } else {
                k_cg_vec<<<vgrid, kT, 0, st>>>(kUpdate, kk, 0, b, scale, x, B);
                ctx->launches += 1;
                r2 = precondition(kUpdate, kk, 0, skip);
            }
            if (r2 != ICQB_OK) return r2;
        }
        ICQB_CUDA(ctx, cudaGetLastError());
        ICQB_CUDA(ctx, cudaMemcpyAsync(&hslot[slot], B.state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
        ICQB_CUDA(ctx, cudaEventRecord(evb[slot], st));
        return ICQB_OK;
    };

#define CG_TRY(expr)                       \
    do {                                   \
        status = (expr);                   \
        if (status != ICQB_OK) goto done;  \
    } while (0)
#define CG_CUDA(expr) CG_TRY(check_cuda(ctx, (expr), #expr))

    {
        CG_CUDA(cudaEventRecord(evTotal0, st));
        k_zero<<<128, 256, 0, st>>>(ctx->d_work, small);
        ctx->launches += 1;
        if (storage == 1) {
            k_recip<<<128, 256, 0, st>>>(ctx->d_area2, inv_area, n);
            ctx->launches += 1;
        }
        // table of the partition: tile -> dense block (every tile belongs to one; blocks of a
        // single tile are dense blocks like the others, so the 128 x 128 set of cg.cu is not built)
        CG_CUDA(cudaMemcpyAsync(dtile, htile.data(), sizeof(int) * (size_t)nT, cudaMemcpyHostToDevice, st));
        if (overlap) {
            CG_CUDA(cudaMemcpyAsync(updBlock, hUpdBlock.data(), sizeof(int) * (size_t)nUpd, cudaMemcpyHostToDevice,
                                    st));
            CG_CUDA(cudaMemcpyAsync(dcover, hCover.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
        }

        // tables of the dense blocks, then: extract, all-reduce (replicated ones), block sweep, mirror
        if (sharded)
            CG_CUDA(cudaMemcpyAsync(dgather, hgather.data(), sizeof(GatherDesc) * hgather.size(),
                                    cudaMemcpyHostToDevice, st));
        if (nDense > 0 || repDoubles > 0) {
            if (nDense > 0)
                CG_CUDA(cudaMemcpyAsync(ddesc, hdesc.data(), sizeof(DenseDesc) * (size_t)nDense,
                                        cudaMemcpyHostToDevice, st));
            if (denseDoubles > 0)
                CG_CUDA(cudaMemsetAsync(dense, 0, sizeof(double) * (size_t)denseDoubles, st));
            const dim3 tgrid((unsigned)std::max<long long>(kMaxTiles, 1), (unsigned)std::max<long long>(kMaxTiles, 1),
                             (unsigned)std::max(nDense, 1));
            if (nDense > 0) {
                if (storage == 1)
                    k_dense_extract_lower<<<tgrid, 256, 0, st>>>(A, L, scale, n, ddesc, dense);
                else
                    k_dense_extract_full<<<tgrid, 256, 0, st>>>(A, ld, mv.r0, mv.r1, scale, n, ddesc, dense);
                ctx->launches += 1;
                CG_CUDA(cudaGetLastError());
            }
            // the replicated blocks (behind the owned ones in the buffer) are summed over the
            // ranks: every rank fills what it stores; same count on every rank
            if (repDoubles > 0) CG_TRY(comm_allreduce_sum(ctx, dense + repOff, repDoubles, st));
        }
        if (nDense > 0) {
            CG_CUDA(cudaFuncSetAttribute(k_sweep_pivot, cudaFuncAttributeMaxDynamicSharedMemorySize, kInvSmem));
            const long long mMax = kMaxTiles * kT;
            const unsigned pairs2 = (unsigned)(kMaxTiles * (kMaxTiles - 1));   // pairs x two column halves
            for (long long K = 0; K < kMaxTiles; ++K) {
                k_sweep_pivot<<<dim3(1, 1, (unsigned)nDense), kInvThreads, kInvSmem, st>>>(
                    ddesc, dense, K, &B.state->weak_pivots, &B.state->bad_blocks);
                k_sweep_panel<<<dim3((unsigned)(mMax / kSwN), 1, (unsigned)nDense), 128, 0, st>>>(ddesc, dense, panels,
                                                                                                K);
                if (pairs2 > 0)
                    k_sweep_update<<<dim3(pairs2, 1, (unsigned)nDense), 128, 0, st>>>(ddesc, dense, panels, K);
                k_sweep_finish<<<dim3((unsigned)kMaxTiles, 1, (unsigned)nDense), 256, 0, st>>>(ddesc, dense, panels, K);
                ctx->launches += 4;
            }
            k_dense_mirror<<<dim3((unsigned)(mMax / kT), (unsigned)mMax, (unsigned)nDense), kT, 0, st>>>(ddesc, dense);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
        }

        // float32 copy of the stored tiles for the late products (mixed precision)
        const bool mixed = (ctx->mixed_switch > 0.0) && storage == 1;
        if (mixed) {
            const long long count = L.numTiles() * kLowerTileElems;
            if (count > ctx->a32_floats) {
                dev_free(ctx, ctx->d_A32);
                ctx->d_A32 = nullptr;
                ctx->a32_floats = 0;
                CG_CUDA(dev_alloc(ctx, &ctx->d_A32, sizeof(float) * (size_t)std::max<long long>(count, 4)));
                ctx->a32_floats = count;
            }
            CG_TRY(launch_f64_to_f32(ctx, A, ctx->d_A32, count, st));
        }

        // z = s * (A x0); r, w, rho_0 and the norms of b
        CG_TRY(mv.apply(x, B.z, scale, st));
        k_cg_vec<<<vgrid, kT, 0, st>>>(kInit, 0, 0, b, scale, x, atIter(0));
        ctx->launches += 1;
        CG_TRY(precondition(kInit, 0, 0, nullptr));

        CgState hs;
        CG_TRY(read_state(ctx, B.state, &hs));
        const bool use_max = (rel == 2);
        const double thr = use_max ? tol * hs.bmax : (rel ? tol * sqrt(hs.bb) : tol);
        const double thr_dev = use_max ? thr : thr * thr;
        double err = use_max ? hs.resmax : sqrt(hs.res2);
        long long k = 0;
        int replacements = 0;
        double true2 = sqrt(hs.res2), trueinf = hs.resmax, prev_true = -1.0;
        bool breakdown = false;
        k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 1);
        k_cg_mixed<<<1, 1, 0, st>>>(B.state, mixed ? ctx->mixed_switch * hs.bmax : 0.0);
        ctx->launches += 2;
        // iterations enqueued per host check, two batches in flight.  Behind the stop flag the
        // kernels of the remaining ones return at once but the collectives of several ranks
        // still run: smaller batches there (at most 7 wasted iterations instead of 15).
        const long long batch = (P > 1) ? 4 : 8;

        while (true) {
            if (err <= thr || k >= maxit) {
                CG_TRY(mv.apply(x, B.z, nullptr, st));
                const bool last = (k >= maxit) || replacements >= 8;
                const int kk = (int)(k & 0x3fffffff);
                k_cg_vec<<<vgrid, kT, 0, st>>>(kTrue, kk, last ? 0 : 1, b, scale, x, atIter(kk));
                ctx->launches += 1;
                CG_TRY(precondition(kTrue, kk, last ? 0 : 1, nullptr));
                CG_TRY(read_state(ctx, B.state, &hs));
                true2 = sqrt(hs.tru2);
                trueinf = hs.trumax;
                const double truem = use_max ? trueinf : true2;
                if (truem <= thr || last) break;
                if (prev_true >= 0.0 && truem > 0.5 * prev_true) break;
                prev_true = truem;
                ++replacements;
                err = truem;
                k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 0);
                k_cg_mixed<<<1, 1, 0, st>>>(B.state, 0.0);   // continue in double precision only
                ctx->launches += 2;
                products32 += hs.n32;
            }
            long long enq = k;
            int inflight = 0, head = 0;
            bool stop = false;
            while (true) {
                while (!stop && inflight < 2 && enq < maxit) {
                    const long long nb = std::min<long long>(batch, maxit - enq);
                    CG_TRY(enqueue_batch(enq, nb, (head + inflight) & 1));
                    enq += nb;
                    ++inflight;
                }
                if (inflight == 0) break;
                CG_CUDA(cudaEventSynchronize(evb[head]));
                hs = hslot[head];
                float gms = 0;
                if (cudaEventElapsedTime(&gms, evg0[head], evg1[head]) == cudaSuccess) {
                    gemv_ms += gms;
                    ++gemv_count;
                }
                head ^= 1;
                --inflight;
                k = hs.iters;
                err = use_max ? hs.resmax : sqrt(hs.res2);
                if (!(hs.res2 == hs.res2)) breakdown = true;
                if (hs.done || breakdown) stop = true;
            }
            if (breakdown) {
                true2 = trueinf = nan("");
                break;
            }
        }
        CG_CUDA(cudaEventRecord(evTotal1, st));
        CG_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, evTotal0, evTotal1);
        ctx->last_ms[kCg] = ms;
        if (gemv_count) ctx->last_ms[kGemv] = gemv_ms / (double)gemv_count;
        ctx->cg_replacements = replacements;
        ctx->cg_products = mv.products;
        ctx->cg_products32 = products32 + hs.n32;
        ctx->cg_bad_blocks = hs.bad_blocks;
        ctx->cg_weak_pivots = hs.weak_pivots;
        if (iters) *iters = k;
        if (resid2) *resid2 = true2;
        if (residinf) *residinf = trueinf;
    }
done:
    if (status != ICQB_OK) cudaStreamSynchronize(st);
    for (int q = 0; q < 2; ++q) {
        cudaEventDestroy(evg0[q]);
        cudaEventDestroy(evg1[q]);
        cudaEventDestroy(evb[q]);
    }
    return status;
#undef CG_TRY
#undef CG_CUDA
}

}  // namespace icqb
