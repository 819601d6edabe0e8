// Preconditioned conjugate gradient on a dense fp64 matrix distributed over the ranks
// of the communicator (row blocks, or the block-lower-triangular storage of the Green
// matrix).
//
// Follows solvers/icqConjugateGradient.py:49-79 (same recurrences, preconditioner
// default diag(A), absolute 2-norm stopping test) with these deliberate differences,
// all documented in DESIGN.md:
//   * optional row scaling s: the iteration runs on diag(s) A x = diag(s) b.  For
//     the Green matrix s = |db x dc| makes the operator symmetric
//     (G[j,i] = G[i,j] A_i/A_j, bem/icqLaplaceMatrices.cpp:97-98); the reference
//     class applied to the raw G does not converge;
//   * the reference recomputes the true residual with a second GEMV every
//     iteration (:74); here the recurrence residual drives the test and the true
//     residual b - A x is recomputed only when the recurrence says "converged"
//     (residual replacement if they disagree) -- one GEMV per iteration;
//   * optional block preconditioner (icqb_cg_set_preconditioner): the exact inverses
//     of the 128 x 128 diagonal blocks of diag(s) A instead of 1/diag -- consecutive
//     triangles are neighbours on the surface, and on the refined spheres this cuts the
//     iteration count 2-3x (131 instead of 306 at 9,800 triangles);
//   * dot products are one partial per CTA, added in a fixed order by the CTA that
//     finishes last (reduce_util.cuh): reproducible and identical on every rank;
//   * the loop state lives on the device (CgState) and iterations are enqueued ahead
//     of the host in batches, two batches in flight, so the GPU never waits for the
//     host between iterations.
//
// Kernels per iteration, lower storage: k_symv (forms p = w + beta p on the fly),
// k_symv_reduce (stores p, z = A p and p.z), k_cg_resid (x, r, w = M^-1 r, rho, residual
// norms, beta, stopping flag).  With several ranks one NCCL all-reduce follows
// k_symv_reduce: it sums the partial products and, in the slot behind them, the ranks'
// p.z_g.  With row-block storage p = w + beta p and p.z are kernels of their own and the
// product's row slices are all-gathered.
#include "cg_kernels.cuh"

using namespace icqb;

extern "C" int icqb_cg_set_preconditioner(icqb_ctx* ctx, int kind) {
    if (!ctx) return ICQB_EARG;
    if (kind < 0 || kind > 2)
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_set_preconditioner: kind must be 0 (diagonal), 1 (blocks) or 2 (blocks + caps)");
    ctx->precond_kind = kind;
    return ICQB_OK;
}

extern "C" int icqb_cg_set_cap_tiles(icqb_ctx* ctx, int64_t head_tiles, int64_t tail_tiles) {
    if (!ctx) return ICQB_EARG;
    if (head_tiles < 0 || tail_tiles < 0 || head_tiles > 64 || tail_tiles > 64)
        return set_error(ctx, ICQB_EARG, "icqb_cg_set_cap_tiles: 0 <= tiles <= 64");
    ctx->cap_head_tiles = head_tiles;
    ctx->cap_tail_tiles = tail_tiles;
    return ICQB_OK;
}

extern "C" int icqb_cg_set_mixed_precision(icqb_ctx* ctx, double switch_rel) {
    if (!ctx) return ICQB_EARG;
    if (!(switch_rel >= 0.0) || switch_rel >= 1.0)
        return set_error(ctx, ICQB_EARG, "icqb_cg_set_mixed_precision: need 0 <= switch_rel < 1");
    ctx->mixed_switch = switch_rel;
    return ICQB_OK;
}

extern "C" int icqb_cg_last_mixed(const icqb_ctx* ctx, int64_t* products32) {
    if (!ctx) return ICQB_EARG;
    if (products32) *products32 = ctx->cg_products32;
    return ICQB_OK;
}

extern "C" int icqb_cg_set_block_partition(icqb_ctx* ctx, const int64_t* starts, int64_t nblocks) {
    if (!ctx) return ICQB_EARG;
    if (nblocks < 0 || (nblocks > 0 && (!starts || starts[0] != 0)))
        return set_error(ctx, ICQB_EARG, "icqb_cg_set_block_partition: starts[0] must be 0");
    for (int64_t q = 1; q < nblocks; ++q)
        if (starts[q] <= starts[q - 1])
            return set_error(ctx, ICQB_EARG, "icqb_cg_set_block_partition: starts must increase");
    ctx->block_starts.assign(starts, starts + nblocks);
    ctx->block_ext0.clear();   // extents belong to a partition: set them after it
    ctx->block_ext1.clear();
    return ICQB_OK;
}

extern "C" int icqb_cg_set_block_extents(icqb_ctx* ctx, const int64_t* ext_begin, const int64_t* ext_end,
                                         int64_t nblocks) {
    if (!ctx) return ICQB_EARG;
    if (nblocks == 0) {
        ctx->block_ext0.clear();
        ctx->block_ext1.clear();
        return ICQB_OK;
    }
    if (nblocks != (int64_t)ctx->block_starts.size() || !ext_begin || !ext_end)
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_set_block_extents: one extent per block of icqb_cg_set_block_partition");
    for (int64_t q = 0; q < nblocks; ++q) {
        const int64_t t0 = ctx->block_starts[(size_t)q];
        if (ext_begin[q] < 0 || ext_begin[q] > t0 || ext_end[q] <= t0 ||
            (q + 1 < nblocks && ext_end[q] < ctx->block_starts[(size_t)q + 1]))
            return set_error(ctx, ICQB_EARG, "icqb_cg_set_block_extents: an extent must contain its block");
        // a row may lie in two extents at most (its own block's and one neighbour's)
        if (q >= 2 && ext_begin[q] < ext_end[q - 2])
            return set_error(ctx, ICQB_EARG, "icqb_cg_set_block_extents: extents of blocks q and q-2 overlap");
    }
    ctx->block_ext0.assign(ext_begin, ext_begin + nblocks);
    ctx->block_ext1.assign(ext_end, ext_end + nblocks);
    return ICQB_OK;
}

extern "C" int icqb_cg_last_blocks(const icqb_ctx* ctx, int64_t* fallback_blocks, int64_t* weak_pivots) {
    if (!ctx) return ICQB_EARG;
    if (fallback_blocks) *fallback_blocks = ctx->cg_bad_blocks;
    if (weak_pivots) *weak_pivots = ctx->cg_weak_pivots;
    return ICQB_OK;
}

extern "C" int icqb_cg_last_stats(const icqb_ctx* ctx, int64_t* replacements, int64_t* products) {
    if (!ctx) return ICQB_EARG;
    if (replacements) *replacements = ctx->cg_replacements;
    if (products) *products = ctx->cg_products;
    return ICQB_OK;
}

extern "C" int icqb_cg_solve_dev(icqb_ctx* ctx, uint64_t dA_, int64_t n, int64_t ld, int storage,
                                 uint64_t drowscale_, uint64_t dprecond_, uint64_t db_,
                                 uint64_t dx_, double tol, int rel, int64_t maxit, int64_t* iters,
                                 double* resid2, double* residinf) {
    if (!ctx) return ICQB_EARG;
    if (!dA_ || !db_ || !dx_ || n <= 0 || (storage == 0 && ld < n))
        return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: null pointer or bad sizes");
    if (storage != 0 && storage != 1)
        return set_error(ctx, ICQB_EARG, "icqb_cg_solve_dev: storage must be 0 (row blocks) or 1 (lower)");
    if (storage == 1 && (ctx->ntri != n || !ctx->d_area2))
        return set_error(ctx, ICQB_ESTATE,
                         "icqb_cg_solve_dev: lower storage needs the mesh of this matrix in the context");
    ICQB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double* A = reinterpret_cast<const double*>(dA_);
    const double* scale = reinterpret_cast<const double*>(drowscale_);
    const double* precond = reinterpret_cast<const double*>(dprecond_);
    const double* b = reinterpret_cast<const double*>(db_);
    double* x = reinterpret_cast<double*>(dx_);
    const int P = comm_nranks(ctx), rank = comm_rank(ctx);
    const long long chunk = (n + P - 1) / P;
    if (maxit < 0) maxit = n;
    cudaStream_t st = ctx->stream;
    if (storage == 1 && scale && scale != ctx->d_area2)
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: with lower storage the row scaling must be icqb_area2_dev()");

    // preconditioner kind 2: larger dense blocks, dense blocks over the sliver tiles (cg_caps.cu)
    if (!precond && ctx->precond_kind == 2)
        return cg_solve_caps(ctx, dA_, n, ld, storage, drowscale_, db_, dx_, tol, rel, maxit, iters,
                             resid2, residinf);

    // work space
    const long long nT = (n + kT - 1) / kT;
    const bool blockPrec = (!precond) && ctx->precond_kind >= 1;
    if (blockPrec && storage == 1 &&
        ((ctx->r1 % kT != 0 && ctx->r1 != n) || (ctx->q1 > ctx->q0 && ctx->q1 % kT != 0 && ctx->q1 != n)))
        return set_error(ctx, ICQB_EARG,
                         "icqb_cg_solve_dev: the block preconditioner needs row blocks that end on "
                         "multiples of 128 (or at n)");
    const long long nfull = chunk * P;
    const long long stateDoubles = (sizeof(CgState) + 7) / 8;
    const long long small = 8 * n + 8 + nfull + 6 * nT + stateDoubles;
    const long long need = small + (blockPrec ? nT * kBlockElems : 0) + 64;
    int rc = ensure_work(ctx, need);
    if (rc != ICQB_OK) return rc;
    CgBuf B;
    double* wp = ctx->d_work;
    B.p = wp; wp += n;
    B.r = wp; wp += n;
    B.w = wp; wp += n;
    B.z = wp; wp += n + 8;   // + the all-reduced p.z slot of the multi-rank lower storage
    B.minv = wp; wp += n;
    B.s = wp; wp += n;
    double* diagA = wp; wp += n;
    double* inv_area = wp; wp += n;
    double* gathered = wp; wp += nfull;
    B.part_rw = wp; wp += nT;
    B.part_res = wp; wp += nT;
    B.part_rmx = wp; wp += nT;
    B.part_a = wp; wp += nT;
    B.part_b = wp; wp += nT;
    B.part_pz = wp; wp += nT;
    B.state = reinterpret_cast<CgState*>(wp); wp += stateDoubles;
    if ((reinterpret_cast<uintptr_t>(wp) & 15) != 0) wp += 1;   // tile_to_regs moves 16-byte pairs
    double* blocks = wp;
    B.Minv = nullptr;
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

    // fusion of the direction update and of p.z into the symmetric product.  With several
    // ranks every rank forms p.(its partial product) and the sum over the ranks rides in the
    // all-reduce of the partial products, one slot behind z: p.z = sum_g p.z_g.
    SymvFuse fuse;
    const bool fuseDir = (storage == 1);
    B.pz = (fuseDir && P > 1) ? B.z + n : &B.state->pz;
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
    CgState* hslot = reinterpret_cast<CgState*>(ctx->h_pinned);   // two read-back slots

    // one batch of iterations kstart .. kstart+nb-1, then the state into read-back slot `slot`
    auto enqueue_batch = [&](long long kstart, long long nb, int slot) -> int {
        for (long long i = 0; i < nb; ++i) {
            const int kk = (int)((kstart + i) & 0x3fffffff);
            if (!fuseDir) {
                k_cg_dir<<<128, 256, 0, st>>>(B);
                ctx->launches += 1;
            }
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg0[slot], st));
            int r2 = mv.apply(B.p, B.z, scale, st, skip, fuseDir ? &fuse : nullptr, fuseDir ? 1 : 0);
            if (r2 != ICQB_OK) return r2;
            if (i == 0) ICQB_CUDA(ctx, cudaEventRecord(evg1[slot], st));
            if (!fuseDir) {
                k_cg_dot<<<vgrid, kT, 0, st>>>(B);
                ctx->launches += 1;
            }
            k_cg_resid<<<vgrid, vblock, 0, st>>>(kUpdate, kk, 0, b, scale, precond, x, B);
            ctx->launches += 1;
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
        LowerLayout L{0, 0, 0, 0, 0, 0, 0};
        if (storage == 1) CG_TRY(lower_layout(ctx, &L));
        if (blockPrec) {
            // inverse diagonal blocks of diag(s) A, replicated on every rank
            if (storage == 1)
                k_cg_blocks_lower<<<vgrid, 256, 0, st>>>(A, L, scale, n, blocks);
            else
                k_cg_blocks_full<<<vgrid, 256, 0, st>>>(A, ld, mv.r0, mv.r1, scale, n, blocks);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
            CG_TRY(comm_allreduce_sum(ctx, blocks, nT * kBlockElems, st));
            CG_CUDA(cudaFuncSetAttribute(k_cg_invert_blocks,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kInvSmem));
            k_cg_invert_blocks<<<vgrid, kInvThreads, kInvSmem, st>>>(blocks, &B.state->bad_blocks);
            ctx->launches += 1;
            CG_CUDA(cudaGetLastError());
            B.Minv = blocks;
        } else if (!precond) {
            // preconditioner: diag(A) unless given
            if (storage == 1) {
                k_cg_diag_lower<<<128, 256, 0, st>>>(A, L, diagA);
                ctx->launches += 1;
            } else if (mv.r1 > mv.r0) {
                k_cg_diag<<<128, 256, 0, st>>>(A, ld, mv.r0, mv.r1, diagA);
                ctx->launches += 1;
            }
            CG_TRY(comm_allreduce_sum(ctx, diagA, n, st));
            precond = diagA;
        }
        // z = s * (A x0); r, w, rho_0 and the norms of b
        CG_TRY(mv.apply(x, B.z, scale, st));
        k_cg_resid<<<vgrid, vblock, 0, st>>>(kInit, 0, 0, b, scale, precond, x, B);
        ctx->launches += 1;
        CG_CUDA(cudaGetLastError());

        CgState hs;
        CG_TRY(read_state(ctx, B.state, &hs));
        // stopping test (rel): 0 absolute 2-norm (the reference class), 1 relative 2-norm,
        // 2 relative max norm  max|b - A x| <= tol * max|b|  (the backward-error criterion)
        const bool use_max = (rel == 2);
        const double thr = use_max ? tol * hs.bmax : (rel ? tol * sqrt(hs.bb) : tol);
        const double thr_dev = use_max ? thr : thr * thr;
        double err = use_max ? hs.resmax : sqrt(hs.res2);
        long long k = 0;
        int replacements = 0;
        double true2 = sqrt(hs.res2), trueinf = hs.resmax, prev_true = -1.0;
        bool breakdown = false;
        k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 1);
        ctx->launches += 1;
        // iterations enqueued per host check, two batches in flight.  Behind the stop flag the
        // kernels of the remaining ones return at once but the collectives of several ranks
        // still run: smaller batches there (at most 7 wasted iterations instead of 15).
        const long long batch = (P > 1) ? 4 : 8;

        while (true) {
            if (err <= thr || k >= maxit) {
                // confirm with the true residual b - A x (and stop, or replace r)
                CG_TRY(mv.apply(x, B.z, nullptr, st));
                const bool last = (k >= maxit) || replacements >= 8;
                k_cg_resid<<<vgrid, vblock, 0, st>>>(kTrue, (int)(k & 0x3fffffff), last ? 0 : 1, b,
                                                     scale, precond, x, B);
                ctx->launches += 1;
                CG_CUDA(cudaGetLastError());
                CG_TRY(read_state(ctx, B.state, &hs));
                true2 = sqrt(hs.tru2);
                trueinf = hs.trumax;
                const double truem = use_max ? trueinf : true2;
                if (truem <= thr || last) break;
                // attainable accuracy reached: the true residual no longer follows the
                // recurrence (it did not even halve since the last replacement)
                if (prev_true >= 0.0 && truem > 0.5 * prev_true) break;
                prev_true = truem;
                ++replacements;
                err = truem;  // r was replaced by the true residual; iterate on
                k_cg_state<<<1, 1, 0, st>>>(B.state, 0, 0, thr_dev, use_max ? 1 : 0, 0);
                ctx->launches += 1;
            }
            // iterations are enqueued `batch` at a time, two batches in flight; the device
            // decides when to stop (CgState), the host only reads the state behind it
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
                if (!(hs.res2 == hs.res2)) breakdown = true;   // NaN
                if (hs.done || breakdown) stop = true;   // later batches are no-ops: drain them
            }
            if (breakdown) {  // singular / indefinite system
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
        ctx->cg_bad_blocks = hs.bad_blocks;
        ctx->cg_weak_pivots = 0;
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
