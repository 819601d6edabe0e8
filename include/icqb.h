/*
 * icqb.h -- C ABI of libicqLaplaceMatricesB200.so, the B200 (sm_100a) replacement
 * for icqsol's boundary-element hot path.
 *
 * The reference's FFI for this path is the ctypes binding in
 * bem/icqBaseLaplaceSolver.py:26-27,68-72 onto
 *     void computeOffDiagonalTerms(vtkPolyData*, double* gMat)   bem/icqLaplaceMatrices.h:6-7
 * plus the quadrature handle API of bem/icqQuadrature.h:19-40.  A vtkPolyData*
 * cannot cross a VTK-free ABI, so the mesh crosses as plain arrays
 * (points double[npts][3], triangles int64[ntri][3]); everything else is plain
 * pointers and sizes.  Device pointers are passed as uint64_t (what
 * torch.Tensor.data_ptr() / cudaMalloc return); streams as uint64_t
 * (cudaStream_t; 0 = the context's own stream).
 *
 * Conventions copied from the reference:
 *   - G is row-major, G[N*obs + src]; "area" = |db x dc| = twice the area;
 *   - off-diagonal entries always use the 16-point (order 8) triangle rule
 *     (bem/icqLaplaceMatrices.cpp:42,94); `diag_order` (1..5) only drives the
 *     self term (bem/icqBaseLaplaceSolver.py:34-66);
 *   - no guards against coincident / zero-area triangles: inf/nan propagate.
 *
 * Every function returns ICQB_OK (0) or a negative status; the message is
 * available from icqb_last_error().  There is no CPU fallback: without a
 * usable CUDA device icqb_create fails.
 */
#ifndef ICQB_H
#define ICQB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICQB_OK 0
#define ICQB_ECUDA (-1)    /* CUDA runtime error                                  */
#define ICQB_EARG (-2)     /* bad argument                                        */
#define ICQB_EORDER (-3)   /* diagonal order outside 1..5 (reference: KeyError,   */
                           /* bem/icqQuadrature1D.py:38)                          */
#define ICQB_ESTATE (-4)   /* call sequence error (no mesh, no matrix ...)        */
#define ICQB_ECOMM (-5)    /* NCCL error / NCCL not loadable                      */

typedef struct icqb_ctx icqb_ctx;

/* ---- context ---------------------------------------------------------------
 * One context per process and GPU (one rank).  Owns the mesh copy, the
 * quadrature-point tables, CG work vectors and (optionally) the NCCL
 * communicator.  Replaces the icqQuadratureInit/Del pair wrapped around the
 * assembly at bem/icqLaplaceMatrices.cpp:40-41,103. */
int icqb_create(icqb_ctx** ctx, int device);
int icqb_destroy(icqb_ctx* ctx);
const char* icqb_last_error(const icqb_ctx* ctx); /* ctx may be NULL: last create error */
int icqb_device(const icqb_ctx* ctx);
uint64_t icqb_stream(const icqb_ctx* ctx);
/* Run all subsequent work of this context on the caller's stream (a cudaStream_t,
 * e.g. torch.cuda.current_stream().cuda_stream; pass 1 = cudaStreamLegacy for the legacy
 * default stream); 0 restores the context's own. */
int icqb_set_stream(icqb_ctx* ctx, uint64_t stream);
int icqb_synchronize(icqb_ctx* ctx);
/* The contexts' device buffers (mesh tables, work space, product plans) come from cudaMalloc
 * behind a process-wide cache of freed blocks, so that building one solver after the other does
 * not pay the driver's map/unmap each time (at most ICQB_ALLOC_CACHE_MB megabytes are kept,
 * default 2048; 0 turns the cache off).  This call hands the cached blocks of `device` (-1: of
 * every device) back to the driver.  No reference counterpart (host malloc/free there). */
int icqb_release_cached_memory(int device);

/* ---- mesh ------------------------------------------------------------------
 * Host arrays, copied.  Replaces the vtkPoints/vtkCellArray traversal at
 * bem/icqLaplaceMatrices.cpp:54-71 (connectivity) and :81-91 (GetPoint).
 * Also precomputes, per triangle, the 16 order-8 quadrature points with the
 * reference's rounding (pa + xi*db + eta*dc, bem/icqQuadrature.cpp:233-239),
 * |db x dc| and the unit normal. */
int icqb_set_mesh(icqb_ctx* ctx, const double* xyz, int64_t npts, const int64_t* tri,
                  int64_t ntri);
int64_t icqb_num_triangles(const icqb_ctx* ctx);
/* observer rows [r0, r1) this context assembles / multiplies (default: all). */
int icqb_set_rows(icqb_ctx* ctx, int64_t r0, int64_t r1);
int icqb_get_rows(const icqb_ctx* ctx, int64_t* r0, int64_t* r1);
/* Block-lower-triangular storage ("lower"): this context owns the
 * observer rows [r0, r1) and [q0, q1) (second block may be empty; blocks start on
 * multiples of 128) and stores, for each row, only the 128-wide column tiles up to and
 * including the diagonal tile.  The upper part follows from the reference's own mirror
 * rule G[j,i] = G[i,j] * area_i / area_j (bem/icqLaplaceMatrices.cpp:97-98). */
int icqb_set_row_blocks(icqb_ctx* ctx, int64_t r0, int64_t r1, int64_t q0, int64_t q1);
/* |db x dc| of every triangle (device pointer, double[ntri]). */
uint64_t icqb_area2_dev(const icqb_ctx* ctx);
int icqb_area2_host(icqb_ctx* ctx, double* out); /* same values, host double[ntri] */

/* ---- uniform refinement ------------------------------------------------------------
 * Every triangle -> k*k children (k per edge), parent orientation kept; the part of the
 * reference's RefineSurface (shapes/icqRefineSurface.py:46-177, cell data copied from the
 * parent to its children at :155-174) that does not need pytriangle.  Output sizes:
 * xyz_out double[ntri*(k+1)(k+2)/2][3] (points duplicated per parent), tri_out
 * int64[ntri*k*k][3]; children of parent t are tri_out[t*k*k .. (t+1)*k*k), so a cell
 * array is inherited by repeating each value k*k times.  float32_points != 0 rounds the new
 * coordinates to float32 (VTK's default point type, shapes/icqShapeManager.py:794-798).
 * Point/child order and roundings equal icqsol_b200.mesh.generators.subdivide (bit-exact).
 * The _dev form takes device pointers and is asynchronous on the context stream. */
int icqb_subdivide_dev(icqb_ctx* ctx, uint64_t dxyz, uint64_t dtri, int64_t ntri, int k,
                       int float32_points, uint64_t dxyz_out, uint64_t dtri_out);
int icqb_subdivide_host(icqb_ctx* ctx, const double* xyz, int64_t npts, const int64_t* tri,
                        int64_t ntri, int k, int float32_points, double* xyz_out, int64_t* tri_out);

/* ---- assembly ----------------------------------------------------------------
 * Fills rows [r0, r1) of G -- and of K when dK != 0 -- into caller-owned device
 * buffers: dG[(r - r0)*ld + c], c in [0, ntri), ld >= ntri.
 *   off-diagonal: bem/icqLaplaceMatrices.cpp:75-101 + bem/icqQuadrature.cpp:208-246
 *                 + bem/icqLaplaceFunctor.cpp:4-10, one fused kernel;
 *   diagonal:     bem/icqBaseLaplaceSolver.py:34-66 (+ icqPotentialIntegrals.py,
 *                 icqReferenceTriangle.py, icqQuadrature1D.py), `diag_order` in 1..5;
 *                 diag_order == 0 leaves the diagonal untouched (the contract of
 *                 computeOffDiagonalTerms itself).
 * K is the double-layer matrix (NOT in the reference snapshot; definition in
 * DESIGN.md); K[i,i] = 0.  Asynchronous on the context stream. */
int icqb_assemble_dev(icqb_ctx* ctx, int diag_order, uint64_t dG, uint64_t dK, int64_t ld);

/* Lower storage, tile-contiguous: the 128-row strips of the two blocks are stored one
 * after the other; a strip with global tile row I holds its tiles J = 0..I back to back,
 * each tile a contiguous row-major 128 x 128 block.  icqb_lower_num_tiles() tiles of
 * 16384 doubles per buffer; tile (strip lt, J) is number tileBase(lt) + J with
 * tileBase(lt) = sum over earlier strips of (their I + 1).  Every pair of triangles is
 * evaluated exactly once over the whole set of ranks and nothing is mirrored.
 * dKl[i,j] = K[i,j] and dKu[i,j] = K[j,i] at the stored positions (both 0 or both set).
 * Rows/columns of ragged edge tiles beyond ntri are never written (zero the buffers). */
int64_t icqb_lower_num_tiles(icqb_ctx* ctx);
int icqb_assemble_lower_dev(icqb_ctx* ctx, int diag_order, uint64_t dG, uint64_t dKl, uint64_t dKu);
/* y = G x (scaled == 0) or diag(area) G x (scaled != 0) from the lower storage: streams
 * each stored entry once (4 N^2 bytes instead of 8 N^2) and all-reduces the partial
 * vectors over the ranks.  x, y: device double[ntri], replicated.  Blocking. */
int icqb_symv_dev(icqb_ctx* ctx, uint64_t dA, uint64_t dx, uint64_t dy, int scaled);
/* The same product on a float32 copy of the stored tiles (made by the call; the timed part is
 * the product alone) -- the kernel behind the optional mixed-precision CG
 * (icqb_cg_set_mixed_precision); the result is G32 x with all sums in double.  B200: 4.6 TB/s
 * of float32 entries (tests/test_gpu_parity.py::test_float32_product_kernel). */
int icqb_symv_f32_dev(icqb_ctx* ctx, uint64_t dA, uint64_t dx, uint64_t dy, int scaled);

/* Host-buffer forms (blocking; host<->device copies inside the call).
 * gMat/kMat: row-major double[(r1-r0)*ntri]; kMat may be NULL. */
int icqb_assemble_host(icqb_ctx* ctx, int diag_order, double* gMat, double* kMat);
/* self terms only: double[r1-r0] (bem/icqBaseLaplaceSolver.py:34-66). */
int icqb_diagonal_host(icqb_ctx* ctx, int diag_order, double* diag);

/* Drop-in for computeOffDiagonalTerms (bem/icqLaplaceMatrices.h:6-7) with the
 * polydata replaced by arrays: fills the off-diagonal of the caller's row-major
 * ntri x ntri host matrix and leaves its diagonal untouched.  Returns status
 * (the reference returns void). */
int computeOffDiagonalTermsArrays(const double* xyz, int64_t npts, const int64_t* tri,
                                  int64_t ntri, double* gMat);

/* ---- dense apply / solve ---------------------------------------------------------
 * y = A x for a row-major device matrix (numpy.dot at bem/icqPoissonSolver.py:36,
 * solvers/icqConjugateGradient.py:57,66,74).  If drowscale != 0, y[i] *= rowscale[i].
 * Bandwidth-bound streaming kernel; asynchronous on `stream`. */
int icqb_gemv_dev(icqb_ctx* ctx, uint64_t dA, int64_t nrows, int64_t ncols, int64_t ld,
                  uint64_t dx, uint64_t dy, uint64_t drowscale, uint64_t stream);
/* Host-vector form on a device matrix: x double[ncols] -> y double[nrows]. */
int icqb_gemv_host(icqb_ctx* ctx, uint64_t dA, int64_t nrows, int64_t ncols, int64_t ld,
                   const double* x, double* y);

/* Arithmetic of the off-diagonal assembly (bem/icqLaplaceMatrices.cpp:75-101,
 * bem/icqQuadrature.cpp:232-243).
 *   variant 0  direct (what a new context uses): every evaluation from the reference's rounded
 *              quadrature points, 12 (G) / 16 (G+K) FP64 instructions per evaluation;
 *   variant 1  (what the solver classes and bench.py select) far-field split: pairs of
 *              triangles with |c_i - c_j| >= 3 (rho_i + rho_j)
 *              (centroids, radii) use the affine form of the observer points about the
 *              observer's own vertex, 9 / 13 instructions per evaluation, entries within a
 *              few ulp of variant 0 (1.2e-15 on a B200, tests/test_gpu_parity.py); all other
 *              pairs are evaluated as in variant 0.  9 % faster at config C2 (G+K).
 * (Round 1's variants 2-4 -- a Newton step in the reciprocal square root, three CTAs per SM --
 * were measured on a B200 in round 2 and removed: the Newton step left 3.3e-13, too close to the
 * 1e-12 contract, and three CTAs per SM were slower.)
 * The stored layouts and every other entry point are unchanged. */
int icqb_set_assembly_variant(icqb_ctx* ctx, int variant);
/* After an assembly with variant 1: how many (warp of 32 observers, source triangle)
 * pairs took the far-field arithmetic, out of how many; zeros with variant 0. */
int icqb_assembly_stats(icqb_ctx* ctx, int64_t* fast_pairs, int64_t* all_pairs);

/* Preconditioned conjugate gradient (solvers/icqConjugateGradient.py:49-79)
 * on the system  diag(rowscale) A x = diag(rowscale) b.
 *   dA       the locally stored part of the n x n matrix (ld is used by storage 0 only):
 *              storage 0: rank g of P holds rows [g*ceil(n/P), (g+1)*ceil(n/P)), all columns;
 *              storage 1: lower storage of the context's row blocks (icqb_set_row_blocks),
 *                         the context must hold the mesh the matrix was assembled from;
 *   drowscale double[n] or 0 (no scaling: plain CG on A, as the reference class);
 *            for the Green matrix pass icqb_area2_dev(): G is not symmetric
 *            (G[j,i] = G[i,j] A_i/A_j, bem/icqLaplaceMatrices.cpp:97-98) but diag(area) G is;
 *            storage 1 requires exactly that vector;
 *   dprecond double[n] diagonal preconditioner of the UNSCALED matrix
 *            (icqConjugateGradient.py:19) or 0 = the context's default
 *            (icqb_cg_set_preconditioner: diag(A), or inverse diagonal blocks);
 *   db, dx   double[n], replicated on every rank; dx holds x0 on entry, the solution on exit;
 *   tol, rel stopping test on the residual b - A x:
 *              rel 0: ||.||_2 <= tol (absolute, icqConjugateGradient.py:64);
 *              rel 1: ||.||_2 <= tol*||b||_2;
 *              rel 2: max|.| <= tol*max|b| (the backward-error criterion of the parity tests);
 *            the recurrence residual drives the loop, the true residual confirms it; if the
 *            true residual stops improving (attainable accuracy) the solve ends early and
 *            the returned residuals say so;
 *   maxit    iteration cap (:16 defaults to n; pass -1 for n);
 *   outputs  iters, final true residual ||b - A x||_2 and max norm.
 * All vectors are replicated and every rank runs the same vector kernels; only the
 * matrix product is distributed: one NCCL collective per iteration (all-gather of the
 * product's row slices for storage 0, all-reduce of the partial products for storage 1).
 * Blocking. */
int icqb_cg_solve_dev(icqb_ctx* ctx, uint64_t dA, int64_t n, int64_t ld, int storage,
                      uint64_t drowscale, uint64_t dprecond, uint64_t db, uint64_t dx, double tol,
                      int rel, int64_t maxit, int64_t* iters, double* resid2, double* residinf);

/* Preconditioner used by icqb_cg_solve_dev when dprecond == 0:
 *   kind 0  diag(A), the reference class's default (solvers/icqConjugateGradient.py:19);
 *   kind 1  the exact inverses of the 128 x 128 diagonal blocks of diag(rowscale) A
 *           (extracted from the matrix, all-reduced over the ranks, inverted and applied on
 *           the device).  Consecutive triangles are neighbours on the surface, so the blocks
 *           hold the strongest couplings: 2-3x fewer iterations on the refined spheres.
 *   kind 2  as kind 1, with the tile rows partitioned into larger diagonal blocks
 *           (icqb_cg_set_block_partition / icqb_cg_set_cap_tiles below): what the solver classes
 *           of the Python layer select (dense blocks over the runs of sliver tiles, blocks of 8
 *           tiles elsewhere).
 * A new context uses kind 0. */
int icqb_cg_set_preconditioner(icqb_ctx* ctx, int kind);
/* Kind 2: as kind 1, but the first `head_tiles` and the last `tail_tiles` 128-row tiles each
 * form ONE dense diagonal block.  Neighbouring sliver triangles make the reference's
 * diag(area) G indefinite and the wrong-sign eigenvectors live on those triangles (DESIGN.md
 * section 4.6); a dense block over them restores the convergence of the CG (B200: the
 * 100,488- and 199,808-triangle UV spheres converge in 171 / 273 iterations; with 128 x 128
 * blocks they do not converge at all).  With kind 2 and no cap tiles the solve is the one of
 * kind 1. */
int icqb_cg_set_cap_tiles(icqb_ctx* ctx, int64_t head_tiles, int64_t tail_tiles);
/* Kind 2, the general form -- diagonal block q of the preconditioner covers the 128-row tiles
 * [starts[q], starts[q+1]) (the last one up to the end), starts[0] = 0, at most 64 tiles per
 * block; blocks of one tile use the 128 x 128 inverses of kind 1, larger ones are extracted as
 * dense matrices, inverted by a symmetric block sweep on the device (csrc/cg_caps.cu) and
 * applied by a batched product.  Blocks of 8 tiles need 91 iterations instead of 165 at config
 * C2 (B200).  With several ranks, blocks that end at the ranks' row-block boundaries are
 * extracted, inverted and applied by the rank that holds their rows.  nblocks = 0 clears the
 * partition (icqb_cg_set_cap_tiles then applies). */
int icqb_cg_set_block_partition(icqb_ctx* ctx, const int64_t* starts, int64_t nblocks);
/* Kind 2, one rank: OVERLAPPING blocks.  Block q of the partition is extended to the tile rows
 * [ext_begin[q], ext_end[q]) (which must contain it; a row may lie in the extents of two
 * consecutive blocks at most) and the preconditioner becomes the symmetric additive Schwarz
 * operator  M^-1 = sum_q R_q^T D_q A_q^-1 D_q R_q,  A_q = the diagonal block of diag(rowscale) A
 * over the extent, D_q = 1/sqrt(number of extents that hold the row).  One tile of overlap on
 * either side of 4-tile blocks halves the iterations on the UV spheres (config C2 on a B200: 45
 * instead of 91, solve 18.3 instead of 32.6 ms; model: tools/experiments/pcg_overlap.py) for the
 * same set-up work and 11 % more traffic in the preconditioner step.  What the solver classes
 * select on one rank (icqRowPartition.blockExtents keeps the runs of sliver tiles whole).  Ignored with several ranks (blocks are then applied by their owners,
 * without overlap).  Call after icqb_cg_set_block_partition; nblocks = 0 clears the extents.
 * No reference counterpart (the reference solves by LU). */
int icqb_cg_set_block_extents(icqb_ctx* ctx, const int64_t* ext_begin, const int64_t* ext_end,
                              int64_t nblocks);
/* Statistics of the last icqb_cg_solve_dev: how often the recurrence residual was replaced
 * by the true residual b - A x (0 = the first confirmation passed), and how many matrix
 * products were enqueued (iterations + confirmations + the ones skipped behind the stop
 * flag). */
int icqb_cg_last_stats(const icqb_ctx* ctx, int64_t* replacements, int64_t* products);
/* Optional, off by default (kind 2, lower storage): mixed-precision products.  Once the
 * recurrence residual max|b - A x| falls below switch_rel * max|b| the products of the
 * iteration stream a float32 copy of the stored tiles (half the bytes; converted once per
 * solve, 2 N^2 more bytes of HBM over all ranks) -- a direction that small tolerates a product
 * accurate to 6e-8 (inexact Krylov); the solve still ends with the double-precision residual
 * b - A x, and if that misses the tolerance it continues in double precision only.  In the
 * numpy model switch_rel = 1e-6 saves 25-28 % of the matrix traffic at an unchanged 5e-13
 * backward error (GPU test: tests/test_gpu_parity.py::test_mixed_precision_cg).  0 (default)
 * = off: the bench and the solver classes compute in double precision throughout. */
int icqb_cg_set_mixed_precision(icqb_ctx* ctx, double switch_rel);
/* Preconditioner set-up of the last solve: 128 x 128 blocks whose inversion met a pivot that was
 * not finite or had lost 13 digits (replaced by 1/diag), and weak pivots met while inverting the
 * dense multi-tile blocks (those are kept; reported only). */
int icqb_cg_last_blocks(const icqb_ctx* ctx, int64_t* fallback_blocks, int64_t* weak_pivots);
/* float32 products of the last solve */
int icqb_cg_last_mixed(const icqb_ctx* ctx, int64_t* products32);

/* ---- direct solve: LU with partial pivoting -----------------------------------------
 * Replaces rsp = numpy.linalg.solve(gMat, src) (bem/icqLaplaceSolver.py:36, LAPACK dgesv) as
 * the solve behind the conjugate gradient that cannot "not converge" (sliver meshes).
 * The matrix is distributed over the ranks of the communicator by COLUMN blocks of 128,
 * block-cyclic: global column block kb lives on rank kb % P as local block kb / P; the local
 * columns are stored one after the other, each contiguous with pitch ld (even, >= n), 16-byte
 * aligned.  icqb_lu_local_columns = how many columns that is for this rank.
 * icqb_lu_factor_dev overwrites dW with the factors; `dscale` (device, length n, may be 0)
 * multiplies local column c by scale[global column of c] first -- the host layer stores ROWS of
 * G there (what the assembly kernels produce) and passes s = |db x dc|, so that the columns are
 * those of the symmetric matrix S = diag(s) G.  *info = 0, or the (1-based) index of the first
 * exactly zero pivot (dgetrf's info > 0).
 * icqb_lu_solve_dev: dx = inv(M) (dscale .* db) for the factored matrix M, i.e. with the same
 * dscale the solution of G x = b; db and dx are full-length device vectors on every rank (dx is
 * the same on every rank afterwards).  Milliseconds: icqb_last_ms(ctx, 5) factorisation,
 * icqb_last_ms(ctx, 6) solve. */
int64_t icqb_lu_local_columns(const icqb_ctx* ctx, int64_t n);
int icqb_lu_factor_dev(icqb_ctx* ctx, uint64_t dW, int64_t n, int64_t ld, uint64_t dscale, int* info);
int icqb_lu_solve_dev(icqb_ctx* ctx, uint64_t dW, int64_t n, int64_t ld, uint64_t dscale,
                      uint64_t db, uint64_t dx);

/* ---- multi-GPU plumbing -----------------------------------------------------------
 * One rank per process.  `unique_id` is the 128-byte ncclUniqueId produced by
 * icqb_comm_unique_id on rank 0 and broadcast by the caller (torch.distributed
 * in the Python host layer). */
int icqb_comm_unique_id(void* unique_id_128);
int icqb_comm_init(icqb_ctx* ctx, const void* unique_id_128, int rank, int nranks);
/* Reuse the communicator of another context of the same process and device
 * (communicator creation costs ~1 s; solvers are created many times). */
int icqb_comm_share(icqb_ctx* ctx, icqb_ctx* owner);
int icqb_comm_rank(const icqb_ctx* ctx, int* rank, int* nranks);
/* Optional peer exchange for the CG product of the lower storage (validated on 4 and 8 B200
 * in round 2: same iterations and solutions as the NCCL path, 0.4 % faster -- so the NCCL
 * all-reduce stays the default; the Python layer attaches it with ICQB_PEER_EXCHANGE=1).  Each rank
 * allocates an exchange buffer for vectors of up to max_n entries (icqb_peer_create returns
 * its 64-byte cudaIpcMemHandle_t), the caller all-gathers the handles and passes the
 * nranks*64 bytes to icqb_peer_attach on every rank.  The solver of preconditioner kind 2
 * then replaces the per-iteration NCCL all-reduce by stores into the peers' buffers from
 * k_symv_reduce and a gather kernel (csrc/cg_caps.cu). */
int icqb_peer_create(icqb_ctx* ctx, int64_t max_n, void* handle64);
int icqb_peer_attach(icqb_ctx* ctx, const void* handles);

/* ---- instrumentation ----------------------------------------------------------------
 * Milliseconds (CUDA events on the context stream) of the last call of each phase:
 * which = 0 prepare, 1 off-diagonal, 2 diagonal, 3 last gemv, 4 last cg solve, 5 LU
 * factorisation, 6 LU solve sweeps;
 * icqb_launch_count: kernels launched by this context since creation. */
double icqb_last_ms(const icqb_ctx* ctx, int which);
int64_t icqb_launch_count(const icqb_ctx* ctx);
/* Sustained FP64 FMA rate of the device, TFLOP/s (register-resident DFMA loop):
 * the measured denominator of the assembly roofline. */
int icqb_measure_fp64_peak(icqb_ctx* ctx, double* tflops);

/* ---- quadrature handle API (bem/icqQuadrature.h:19-40) --------------------------------
 * Same names, argument order and meaning as the reference; each Evaluate call
 * runs the device code on one CUDA block so that testQuadratureCpp-style
 * known-answer tests exercise the GPU arithmetic.  icqQuadratureSetFunctor takes
 * a C++ object in the reference and is not reproducible across a C ABI: omitted
 * (the functor is always the Laplace kernel of bem/icqLaplaceFunctor.cpp:4-10). */
typedef struct icqQuadratureType icqQuadratureType;
void icqQuadratureInit(icqQuadratureType** self);
void icqQuadratureSetObserver(icqQuadratureType** self, const double* pObs);
int icqQuadratureGetMaxOrder(icqQuadratureType** self);
double icqQuadratureEvaluate(icqQuadratureType** self, int order, const double* pa,
                             const double* pb, const double* pc);
double icqQuadratureEvaluateDouble(icqQuadratureType** self, int order, const double* pa0,
                                   const double* pb0, const double* pc0, const double* pa1,
                                   const double* pb1, const double* pc1);
void icqQuadratureDel(icqQuadratureType** self);

#ifdef __cplusplus
}
#endif
#endif /* ICQB_H */
