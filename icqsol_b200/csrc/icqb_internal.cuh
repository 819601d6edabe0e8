// Internal declarations shared by the translation units of
// libicqLaplaceMatricesB200.so.  The public surface is include/icqb.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/icqb.h"

namespace icqb {

constexpr int kNq = 16;          // nodes of the order-8 triangle rule
constexpr int kQpStride = 48;    // doubles per triangle in the quadrature-point tables
constexpr int kPad = 128;        // triangle count is padded to a multiple of this

// phases for icqb_last_ms
enum Phase { kPrepare = 0, kOffDiag = 1, kDiag = 2, kGemv = 3, kCg = 4, kLu = 5, kLuSolve = 6, kNumPhases = 7 };

// Tile-contiguous layout of the block-lower-triangular storage: the locally owned
// 128-row strips are stored one after the other; strip lt (global tile row I) holds its
// tiles J = 0..I back to back, each tile a contiguous row-major 128 x 128 block (128 KB).
// The stored tiles of a rank are therefore ONE contiguous array, which the symmetric product
// cuts evenly into per-warp spans of 8 KB chunks, and the storage is exactly the lower
// triangle: 4 N^2 + O(128 N) bytes over all ranks.
constexpr int kLowerTile = 128;
constexpr long long kLowerTileElems = (long long)kLowerTile * kLowerTile;

struct LowerLayout {
    long long r0, r1, q0, q1;   // row blocks (non-empty ones start on multiples of 128)
    int nTiles0, nTiles1;       // strips per block
    long long tiles0;           // tiles stored for the first block

    __host__ __device__ int nLocal() const { return nTiles0 + nTiles1; }
    // global tile row of local strip lt
    __host__ __device__ long long tileRow(int lt) const {
        return lt < nTiles0 ? r0 / kLowerTile + lt : q0 / kLowerTile + (lt - nTiles0);
    }
    // index of tile (lt, J = 0) in the local tile sequence
    __host__ __device__ long long tileBase(int lt) const {
        if (lt < nTiles0) {
            const long long I0 = r0 / kLowerTile, k = lt;
            return k * (I0 + 1) + k * (k - 1) / 2;
        }
        const long long I1 = q0 / kLowerTile, k = lt - nTiles0;
        return tiles0 + k * (I1 + 1) + k * (k - 1) / 2;
    }
    __host__ __device__ long long numTiles() const { return tileBase(nLocal()); }
    // first global row, one-past-last row of the block, of strip lt
    __host__ __device__ long long obs0(int lt) const {
        return lt < nTiles0 ? r0 + (long long)lt * kLowerTile
                            : q0 + (long long)(lt - nTiles0) * kLowerTile;
    }
    __host__ __device__ long long obsEnd(int lt) const { return lt < nTiles0 ? r1 : q1; }
    // local strip of global row i, or -1
    __host__ __device__ int stripOf(long long i) const {
        if (i >= r0 && i < r1) return (int)((i - r0) / kLowerTile);
        if (i >= q0 && i < q1) return nTiles0 + (int)((i - q0) / kLowerTile);
        return -1;
    }
};

struct Comm;      // NCCL state, comm.cu
struct SymvPlan;  // task table + workspace of the symmetric product, symv.cu

// Loop state of the conjugate gradient, kept on the device so that iterations can be
// enqueued ahead of the host (cg.cu).  The last CTA of a reducing kernel writes the
// scalars; the next kernel in the stream reads them.
struct CgState {
    int done;                  // recurrence residual met the threshold: later kernels return at once
    int use_max;               // stopping test on the max norm instead of the 2-norm
    unsigned int ticket_dot;   // "last CTA" counters of the two reducing kernels
    unsigned int ticket_res;
    long long iters;           // iterations completed (stops counting once done)
    double thr;                // threshold: squared 2-norm, or max norm (use_max)
    double res2, resmax;       // recurrence residual of the unscaled system
    double rho[2];             // r.w of iteration k at rho[k & 1]
    double pz;                 // p.(A p)
    double beta;               // rho_{k+1}/rho_k for the next direction
    double bb, bmax;           // |b|_2^2, max|b|
    double tru2, trumax;       // true residual |b - A x|_2^2, max norm
    // optional mixed-precision products (cg_caps.cu only; appended: the other kernels'
    // offsets are unchanged): once the recurrence residual max|r/s| falls below switch_thr the
    // products read the float32 copy of the matrix
    int lowp;                  // 1: float32 products
    int skip64, skip32;        // done | lowp,  done | !lowp  (the two matrix kernels' skip flags)
    int n32;                   // float32 products so far
    double switch_thr;         // 0: mixed precision off
    int bad_blocks;            // 128 x 128 preconditioner blocks that fell back to 1/diag (weak pivot)
    int weak_pivots;           // weak pivots met while inverting the dense blocks (reported only)
};

// Conjugate-gradient work fused into the symmetric product (single rank, lower storage):
// the multiplied vector is the new direction p = w + beta * p_old, formed on the fly by
// k_symv and stored by k_symv_reduce, which also leaves p.(A p) in *pz.
struct SymvFuse {
    const double* w = nullptr;       // null: plain product of x
    const double* beta = nullptr;    // device scalar
    double* p = nullptr;             // the vector x itself, writable
    double* pz_part = nullptr;       // [ceil(n/128)] partials; null: no fused dot product
    double* pz = nullptr;
    unsigned int* ticket = nullptr;
};

// Optional peer exchange of the CG product (several ranks, lower storage): instead of an
// NCCL all-reduce, k_symv_reduce stores its partial product and its p.z straight into a slot
// of EVERY rank's exchange buffer (peer memory over NVLink, CUDA IPC) and raises a flag there;
// k_peer_gather on each rank waits for the P flags and adds the P slots in rank order.
// Buffer of a rank: flags [2 parities][P] (256 bytes reserved), then slots
// [2 parities][P][slotStride] doubles; slot (parity, g) is written by rank g only.
constexpr int kMaxPeers = 16;
struct SymvPeer {
    int nranks = 0, rank = 0, parity = 0;
    unsigned long long seq = 0;             // flags are raised to the product's sequence number
    long long slotStride = 0;               // doubles per slot (>= n + 1)
    double* slots[kMaxPeers] = {};          // per rank h: its slot array
    unsigned long long* flags[kMaxPeers] = {};   // per rank h: its flag array
};

}  // namespace icqb

struct icqb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // created with the context
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = 0;

    int64_t npts = 0, ntri = 0, ntri_pad = 0;
    int64_t r0 = 0, r1 = 0;   // observer rows owned by this context (first block)
    int64_t q0 = 0, q1 = 0;   // second row block (block-lower-triangular storage only)

    // mesh and per-triangle tables (device)
    double* d_xyz = nullptr;     // [npts][3]
    int64_t* d_tri = nullptr;    // [ntri][3]
    double* d_verts = nullptr;   // [ntri][9]  gathered vertices a,b,c
    double* d_qp_soa = nullptr;  // [48][ntri_pad]   component-major: observer side
    double* d_qp_aos = nullptr;  // [ntri_pad][48]   triangle-major: source side (bulk-copied tiles)
    double* d_area2 = nullptr;   // [ntri_pad]  |db x dc|
    double* d_nrm = nullptr;     // [ntri_pad][4]  unit normal, n . pa
    // far-field split of the off-diagonal assembly (assemble_offdiag_ff.cu)
    int offdiag_variant = 0;     // 0 direct kernel (k_offdiag), 1 far-field split (k_offdiag_ff: what the
                                 // solver classes and bench.py select)
    double coord_max = 0.0;      // max |coordinate| of the mesh points
    double* d_bound = nullptr;   // [ntri_pad + 128][4]  centroid, radius; built on first use
    unsigned long long* d_ffstats = nullptr;   // [2] (warp, source) pairs on the fast path / all

    // CG work space (device), sized on demand
    double* d_work = nullptr;
    int64_t work_doubles = 0;
    double* h_pinned = nullptr;  // small pinned scratch for scalar read-back

    icqb::Comm* comm = nullptr;
    icqb::SymvPlan* symv = nullptr;

    int precond_kind = 0;   // CG preconditioner when none is passed: 0 diag(A), 1 inverse 128x128 diagonal blocks,
                            // 2 larger diagonal blocks: block_starts, or caps over cap_head_tiles / cap_tail_tiles
    int64_t cap_head_tiles = 0, cap_tail_tiles = 0;
    std::vector<int64_t> block_starts;   // kind 2: first tile row of every diagonal block
    std::vector<int64_t> block_ext0, block_ext1;   // kind 2, one rank: overlapping extents [ext0, ext1) of the blocks
    int64_t cg_replacements = 0, cg_products = 0;   // statistics of the last solve
    int64_t cg_bad_blocks = 0, cg_weak_pivots = 0;  // preconditioner blocks replaced by 1/diag, weak pivots
    double mixed_switch = 0.0;     // optional: relative residual below which products use float32 (0 = off)
    float* d_A32 = nullptr;        // (mixed precision) float32 copy of the stored tiles
    int64_t a32_floats = 0;
    int64_t cg_products32 = 0;     // float32 products of the last solve (from the device state)

    // LU with partial pivoting (lu.cu): pivot rows of the last factorisation, panel buffer, small state
    long long* d_lu_piv = nullptr;
    double* d_lu_buf = nullptr;
    double* d_lu_small = nullptr;
    int64_t lu_n = 0;

    double last_ms[icqb::kNumPhases] = {0, 0, 0, 0, 0, 0, 0};
    int64_t launches = 0;
    std::string err;
};

namespace icqb {

int set_error(icqb_ctx* ctx, int code, const std::string& msg);
// layout of the context's row blocks (sets the context error if a block is misaligned)
int lower_layout(icqb_ctx* ctx, LowerLayout* L);
int check_cuda(icqb_ctx* ctx, cudaError_t e, const char* what);
int ensure_work(icqb_ctx* ctx, int64_t doubles);
// Device buffers of the context (mesh tables, work space, product plan): cudaMalloc behind a
// process-wide cache of freed blocks keyed by (device, size) -- context.cu; a process that builds
// one solver after the other gets its ~20 buffers back from the cache instead of from the driver.
// dev_free waits for the context's stream before it keeps a block.
cudaError_t dev_alloc_raw(icqb_ctx* ctx, void** p, size_t bytes);
void dev_free(icqb_ctx* ctx, void* p);
template <class T>
inline cudaError_t dev_alloc(icqb_ctx* ctx, T** p, size_t bytes) {
    return dev_alloc_raw(ctx, reinterpret_cast<void**>(p), bytes);
}

// kernels' host launchers (each returns ICQB_OK or sets the context error)
int launch_prepare(icqb_ctx* ctx);
int launch_offdiag(icqb_ctx* ctx, double* dG, double* dK, int64_t ld);
int launch_offdiag_lower(icqb_ctx* ctx, double* dG, double* dK, double* dKu);
int launch_diag_rows(icqb_ctx* ctx, int order, long long g0, long long g1, double* dOut,
                     int64_t stride);
int launch_symv(icqb_ctx* ctx, const double* dA, const double* dx,
                const double* dgamma, const double* dalpha, const double* dbeta, double* dz,
                cudaStream_t stream, const int* skip = nullptr, const SymvFuse* fuse = nullptr,
                const SymvPeer* peer = nullptr);
// optional float32 products of the mixed-precision CG (symv.cu)
int launch_symv_mixed(icqb_ctx* ctx, const double* dA, const float* dA32, const double* dx,
                      const double* dgamma, const double* dalpha, const double* dbeta, double* dz,
                      cudaStream_t stream, const int* done, const int* skip64, const int* skip32,
                      const SymvFuse* fuse, const SymvPeer* peer);
int launch_f64_to_f32(icqb_ctx* ctx, const double* dA, float* dA32, long long count,
                      cudaStream_t stream);
int launch_symv_f32_only(icqb_ctx* ctx, const float* dA32, const double* dx, const double* dgamma,
                         const double* dalpha, const double* dbeta, double* dz, cudaStream_t stream);
int launch_diag(icqb_ctx* ctx, int order, double* dOut, int64_t stride);
int launch_gemv(icqb_ctx* ctx, const double* dA, int64_t nrows, int64_t ncols, int64_t ld,
                const double* dx, double* dy, const double* drowscale, cudaStream_t stream,
                const int* skip = nullptr);

// CG with larger dense blocks in the preconditioner (kind 2, cg_caps.cu); arguments as
// icqb_cg_solve_dev without the diagonal preconditioner
int cg_solve_caps(icqb_ctx* ctx, uint64_t dA, int64_t n, int64_t ld, int storage, uint64_t drowscale,
                  uint64_t db, uint64_t dx, double tol, int rel, int64_t maxit, int64_t* iters,
                  double* resid2, double* residinf);

// NCCL plumbing (comm.cu); all no-ops for a single rank
int comm_nranks(const icqb_ctx* ctx);
int comm_rank(const icqb_ctx* ctx);
int comm_allgather(icqb_ctx* ctx, const double* send, double* recv, int64_t count_per_rank,
                   cudaStream_t stream);
int comm_allreduce_sum(icqb_ctx* ctx, double* buf, int64_t count, cudaStream_t stream);
int comm_allreduce_max(icqb_ctx* ctx, double* buf, int64_t count, cudaStream_t stream);
// in place: `buf` of rank `root` to every rank
int comm_broadcast(icqb_ctx* ctx, double* buf, int64_t count, int root, cudaStream_t stream);
void lu_free(icqb_ctx* ctx);
void comm_destroy(icqb_ctx* ctx);
// optional peer exchange (comm.cu): fills `out` for a product over n entries and draws its
// sequence number; false when no exchange is attached or n does not fit
bool comm_peer_next(icqb_ctx* ctx, long long n, SymvPeer* out);
void symv_plan_free(icqb_ctx* ctx);

}  // namespace icqb

#define ICQB_CUDA(ctx, call)                                              \
    do {                                                                  \
        int _st = icqb::check_cuda((ctx), (call), #call);                 \
        if (_st != ICQB_OK) return _st;                                   \
    } while (0)
