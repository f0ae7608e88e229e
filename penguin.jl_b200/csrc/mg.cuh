// mg.cuh -- geometric multigrid preconditioner of the folded CG for systems WITHOUT interface unknowns (monophasic, Dirichlet interface): the
// steady Poisson problem of BASELINE.json configs[4] (src/solver/diffusion.jl:14-72) has no V / dt shift, so kappa(M^) = O(n^2) and the
// block-Jacobi-scaled CG of fold_solve needs ~3.7 n iterations (measured: 459 / 929 / 1404 at 128^3 / 256^3 / 384^3 around 64 spheres).
//
// Design (prototyped on the oracle's matrices: tests/experiments/mg_experiment.py):
//   * levels are REDISCRETISED: the capacity kernels run again on the meshes n/2, n/4, ... for the same level set, and each level is a complete
//     folded system of its own (fold_build on a coarse pb200_solver) -- the coarse operators come out of the same geometry + folding code as the
//     fine one, and every level is applied by the same TMA-staged kernels;
//   * transfers are cell aggregation: a coarse cell is the union of its 2^N children.  The rows of M are integrated over cells, so the residual
//     restricts by SUMMATION (R = P^T) and the correction prolongs by injection.  Both act on the true unknowns / rows: with the diagonal scaling
//     x = S x^, r = S^-1 r^ (S = diag(sc) = 1 / sqrt(diag M)) of the folded system,  r^_c = S_c P^T S_f^-1 r^_f  and  x^_f += S_f^-1 P S_c x^_c;
//   * smoother: the Chebyshev polynomial q_m(M^) of fold_poly (MODE 4 of the apply kernels: the operator apply with an axpby epilogue) on the
//     upper part [lambda_max / alpha, lambda_max] of each level's spectrum, before and after the coarse correction -- a polynomial in M^, hence
//     symmetric, so the V-cycle is a symmetric positive definite preconditioner and CG stays CG;
//   * coarsest level: a fixed number of such sweeps on [lambda_max / alpha_c, lambda_max] (a fixed polynomial again).
// One rank only (the slab decomposition does not coarsen with the grid); constant diffusion coefficient.
#pragma once

struct MgXfer {
    int N, sd;
    long long ldf[3], ldc[3];   // local array extents (ghost planes of the slab dimension included), fine / coarse
    long long P0f, P0c;         // x pitch of the re-pitched Krylov vectors, fine / coarse
    int ncc[3];                 // real coarse cells per dimension
};

// local array coordinates of cell k of this thread in bulk tile R (the arithmetic of tile_cell, without the linear index)
__device__ __forceinline__ void mg_cell_coords(const Items &I, const TileRec &R, int k, int c[3])
{
    const int tx = (int)threadIdx.x & ((1 << I.shx) - 1), ty = (int)threadIdx.x >> I.shx;
    c[0] = R.ox + tx + I.kx * k; c[1] = R.oy + ty * I.tym + I.ky * k; c[2] = R.oz + I.kz * k;
}
// r^_c = S_c sum_children r^_f / S_f over the coarse items (inactive cells carry sc = 0 and contribute / receive nothing)
__global__ void __launch_bounds__(FCH) kf_mg_restrict(Items Ic, MgXfer X, const double *__restrict__ scf, const double *__restrict__ scc, const double *__restrict__ rf,
                                                      double *__restrict__ rc)
{
    FV_LOOP(Ic) {
        if (f != 0) continue;
        int c[3];
        mg_cell_coords(Ic, R__, k__, c);
        const double sc = scc[i];
        bool real = sc > 0.0;
        for (int d = 0; d < X.N; ++d) { if (d == X.sd) c[d] -= 1; if (c[d] < 0 || c[d] >= X.ncc[d]) real = false; }
        double sum = 0.0;
        if (real) {
            const int nz = X.N > 2 ? 2 : 1, ny = X.N > 1 ? 2 : 1;
            for (int cz = 0; cz < nz; ++cz)
                for (int cy = 0; cy < ny; ++cy)
                    for (int cx = 0; cx < 2; ++cx) {
                        const long long xf = 2 * c[0] + cx + (X.sd == 0 ? 1 : 0);
                        const long long yf = X.N > 1 ? 2 * c[1] + cy + (X.sd == 1 ? 1 : 0) : 0;
                        const long long zf = X.N > 2 ? 2 * c[2] + cz + (X.sd == 2 ? 1 : 0) : 0;
                        const long long lf = xf + X.ldf[0] * (yf + X.ldf[1] * zf);
                        const double s = scf[lf];
                        if (s > 0.0) sum += rf[xf + X.P0f * (yf + X.ldf[1] * zf)] / s;
                    }
        }
        rc[q] = sc * sum;
    }
}
// x^_f += S_f^-1 P S_c x^_c over the fine items
__global__ void __launch_bounds__(FCH) kf_mg_prolong(Items If, MgXfer X, const double *__restrict__ scf, const double *__restrict__ scc, const double *__restrict__ ec,
                                                     double *__restrict__ ef)
{
    FV_LOOP(If) {
        if (f != 0) continue;
        const double s = scf[i];
        if (!(s > 0.0)) continue;
        int c[3];
        mg_cell_coords(If, R__, k__, c);
        bool real = true;
        for (int d = 0; d < X.N; ++d) { if (d == X.sd) c[d] -= 1; c[d] >>= 1; if (c[d] < 0 || c[d] >= X.ncc[d]) real = false; }
        if (!real) continue;
        const long long xc = c[0] + (X.sd == 0 ? 1 : 0), yc = X.N > 1 ? c[1] + (X.sd == 1 ? 1 : 0) : 0, zc = X.N > 2 ? c[2] + (X.sd == 2 ? 1 : 0) : 0;
        const long long lc = xc + X.ldc[0] * (yc + X.ldc[1] * zc);
        ef[q] += scc[lc] * ec[xc + X.P0c * (yc + X.ldc[1] * zc)] / s;
    }
}
__global__ void __launch_bounds__(FCH) kf_mg_add(Items I, FVec a, FVec x) { FV_LOOP(I) { (void)i; x.f[f][q] += a.f[f][q]; } }
// CG: alpha = rho / (p, v);  x += alpha p;  r -= alpha v;  publishes (r, r)
__global__ void __launch_bounds__(FCH) kf_mg_xr(Items I, const double *res, int sl_rho, int sl_sig, FVec p, FVec v, FVec x, FVec r, double *partials, double *results,
                                                unsigned *counter)
{
    const double alpha = safe_div(res[sl_rho], res[sl_sig]);
    double s[1] = {0.0};
    FV_LOOP(I) {
        (void)i;
        x.f[f][q] += alpha * p.f[f][q];
        const double rn = r.f[f][q] - alpha * v.f[f][q];
        r.f[f][q] = rn;
        s[0] += rn * rn;
    }
    block_reduce_publish<1>(s, partials, results, counter);
}
// p = z + beta p, beta = rho_new / rho_old (first: p = z)
__global__ void __launch_bounds__(FCH) kf_mg_p(Items I, const double *res, int sl_new, int sl_old, int first, FVec z, FVec p)
{
    const double beta = first ? 0.0 : safe_div(res[sl_new], res[sl_old]);
    FV_LOOP(I) { (void)i; p.f[f][q] = z.f[f][q] + beta * p.f[f][q]; }
}

struct MgLevel {
    pb200_solver *s = nullptr;         // level 0: the caller's solver (not owned)
    pb200_capacity *cap = nullptr;     // coarse levels: owned
    pb200_ops *ops = nullptr;
    FVec r = {}, e = {}, res = {}, tmp = {};
    double lam = 2.0;                  // lambda_max(M^) of the level (power iteration)
    MgXfer X;                          // transfer to the NEXT (coarser) level
};
struct MgHier {
    std::vector<MgLevel> lev;
    bool ready = false;
    cudaGraphExec_t exec = nullptr;    // one MG-PCG iteration (mg_pcg)
    int64_t graph_launches = 0, graph_applies = 0;
    int deg = 2, coarse_sweeps = 12;
    double alpha = 3.0, alpha_c = 40.0;
    double min_res = 1.0;              // coarsen while the smallest ball radius spans at least this many coarse cells (384^3: 1.0 -> 43 iterations, 1.5 -> 55)
};

static void mg_free(pb200_solver *s)
{
    MgHier *H = s->mg;
    if (!H) return;
    for (size_t l = 0; l < H->lev.size(); ++l) {
        MgLevel &L = H->lev[l];
        FVec *vs[] = {&L.r, &L.e, &L.res, &L.tmp};
        for (FVec *v : vs) fold_free_vec(*v);
        if (l > 0) { pb200_solver_destroy(L.s); pb200_ops_destroy(L.ops); pb200_capacity_destroy(L.cap); }
    }
    if (H->exec) cudaGraphExecDestroy(H->exec);
    delete H;
    s->mg = nullptr;
}

// lambda_max(M^) by a power iteration without normalisation (lambda_max < ~2: 2^24 is harmless in fp64), as fold_poly_setup does
static int mg_lambda_max(pb200_solver *s, double *lam)
{
    pb200_ctx *ctx = s->ctx;
    FoldSys &F = s->F;
    int rc;
    const int gz = wave_grid(s, kf_seed);
    kf_seed<<<gz, FCH, 0, ctx->stream>>>(F.I, F.p); LAUNCH_CHECK(ctx);
    FVec *a = &F.p, *b = &F.v;
    for (int it = 0; it < 24; ++it) {
        if ((rc = fold_apply(s, *a, *b, *b, 3))) return rc;
        std::swap(a, b);
    }
    double t2[2];
    if ((rc = fetch_results(ctx, FS_TS_D, 2, t2))) return rc;   // (y, x), (y, y) of the last apply (no band part: has_w is false here)
    *lam = (t2[0] > 0.0 && t2[1] > 0.0) ? t2[1] / t2[0] : 2.0;
    kf_zero<<<gz, FCH, 0, ctx->stream>>>(F.I, F.p); LAUNCH_CHECK(ctx);
    kf_zero<<<gz, FCH, 0, ctx->stream>>>(F.I, F.v); LAUNCH_CHECK(ctx);
    return PB200_OK;
}

// out = q_m(M^) r for the Chebyshev interval [lo, hi] (m = 1, 2; m = 2 goes through F.z)
static int mg_poly(pb200_solver *s, const FVec &r, const FVec &out, double lo, double hi, int m)
{
    PolyCoef c[2];
    poly_coefs(lo, hi, c);
    const StopCrit ns = {0.0, 0.0, -1};
    int rc;
    if (m <= 1) return fold_apply(s, r, out, r, 4, ns, c[0], FS_TMP, FS_TMP + 1);
    if ((rc = fold_apply(s, r, s->F.z, r, 4, ns, c[0], FS_TMP, FS_TMP + 1))) return rc;
    return fold_apply(s, s->F.z, out, r, 4, ns, c[1], FS_TMP, FS_TMP + 1);
}
// res = r - M^ e
static int mg_residual(pb200_solver *s, const FVec &r, const FVec &e, const FVec &res)
{
    return fold_apply(s, e, res, r, 4, StopCrit{0.0, 0.0, -1}, PolyCoef{1.0, 0.0, -1.0}, FS_TMP, FS_TMP + 1);
}

static int build_masks(pb200_solver *s);
static int mg_setup(pb200_solver *s, const ApplyCoef &ac)
{
    pb200_ctx *ctx = s->ctx;
    if (ctx->nranks > 1) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner runs on one rank only");
    if (s->sp.phase_type != PB200_MONO || s->F.d.has_w) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner needs a monophasic system with a Dirichlet interface (no interface unknowns)");
    if (s->g.N < 2) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner needs a 2-D or 3-D grid");
    if (s->D1arr) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner needs a constant diffusion coefficient");
    for (int k = 0; k < 6; ++k)
        if (s->bd.present[k] && s->bd.kind[k] == PB200_BC_PERIODIC) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner does not handle Periodic border rows");
    const pb200_capacity *c0 = s->o1->cap;
    if (!c0->has_ls) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner rebuilds the capacities on coarser meshes: imported capacities carry no level set");
    mg_free(s);
    MgHier *H = new MgHier();
    s->mg = H;
    if (getenv("PB200_MG_DEG")) H->deg = atoi(getenv("PB200_MG_DEG")) >= 2 ? 2 : 1;
    if (getenv("PB200_MG_ALPHA")) H->alpha = atof(getenv("PB200_MG_ALPHA"));
    if (getenv("PB200_MG_SWEEPS")) H->coarse_sweeps = atoi(getenv("PB200_MG_SWEEPS"));
    if (getenv("PB200_MG_ALPHAC")) H->alpha_c = atof(getenv("PB200_MG_ALPHAC"));
    if (!(H->alpha > 1.5)) H->alpha = 3.0;
    if (getenv("PB200_MG_RES")) H->min_res = atof(getenv("PB200_MG_RES"));
    int maxlev = getenv("PB200_MG_LEVELS") ? atoi(getenv("PB200_MG_LEVELS")) : 12;
    H->lev.emplace_back();
    H->lev[0].s = s;
    int rc;
    pb200_levelset ls;
    ls.kind = c0->ls_kind; ls.nballs = (int)c0->ls_r.size(); ls.centers = c0->ls_c.data(); ls.radii = c0->ls_r.data();
    ls.fluid_inside = c0->ls_inside; ls.hs_dim = c0->ls_hsdim; ls.hs_c = c0->ls_hsc;
    while ((int)H->lev.size() < maxlev) {
        const Grid &gf = H->lev.back().s->g;
        bool ok = true;
        int n[3] = {1, 1, 1};
        double hc = 0.0;
        for (int d = 0; d < gf.N; ++d) { if (gf.nc[d] % 2 || gf.nc[d] / 2 < 4) ok = false; n[d] = gf.nc[d] / 2; hc = fmax(hc, 2.0 * gf.h[d]); }
        // A level whose cells are larger than the bodies misrepresents their Dirichlet condition and HURTS the cycle (measured, 256^3 around 64 spheres of
        // radius 0.1-0.3: 3 levels 31 iterations, 4 levels 45, 5 levels 62, 7 levels 75): stop while the smallest ball still spans min_res coarse cells.
        if (ls.kind == PB200_LS_BALLS && !c0->ls_r.empty() && *std::min_element(c0->ls_r.begin(), c0->ls_r.end()) < H->min_res * hc) ok = false;
        if (!ok) break;
        MgLevel L;
        if ((rc = pb200_capacity_create(ctx, gf.N, n, gf.x0, gf.L, &ls, 0, &L.cap))) return rc;
        if ((rc = pb200_ops_create(L.cap, &L.ops))) { pb200_capacity_destroy(L.cap); return rc; }
        pb200_solver_desc d;
        memset(&d, 0, sizeof(d));
        d.phase_type = PB200_MONO; d.time_type = s->sp.time_type; d.ops1 = L.ops; d.D1 = s->p1.Dc; d.D2 = s->p1.Dc; d.ifc_kind = PB200_BC_DIRICHLET;
        if ((rc = pb200_solver_create(ctx, &d, &L.s))) { pb200_ops_destroy(L.ops); pb200_capacity_destroy(L.cap); return rc; }
        H->lev.push_back(L);                                   // (owned from here on: mg_free releases it)
        pb200_solver *sc = H->lev.back().s;
        for (int k = 0; k < 6; ++k)
            if (s->bd.present[k] && (rc = pb200_solver_set_border(sc, k, s->bd.kind[k], 0.0, nullptr))) return rc;
        if ((rc = build_masks(sc)) || (rc = fold_build(sc, ac))) return rc;
        if (sc->F.d.has_w || sc->dof_bulk < 1) { H->lev.pop_back(); pb200_solver_destroy(sc); pb200_ops_destroy(L.ops); pb200_capacity_destroy(L.cap); break; }
    }
    for (size_t l = 0; l < H->lev.size(); ++l) {
        MgLevel &L = H->lev[l];
        pb200_solver *sl = L.s;
        FoldSys &F = sl->F;
        if (!F.have_z) { if ((rc = fold_alloc_vec(sl, &F.z))) return rc; F.have_z = true; }
        FVec *vs[] = {&L.r, &L.e, &L.res, &L.tmp};
        for (FVec *v : vs) if ((l > 0 || v != &L.r) && (rc = fold_alloc_vec(sl, v))) return rc;
        if ((rc = mg_lambda_max(sl, &L.lam))) return rc;
        L.lam *= 1.05;
        if (l + 1 < H->lev.size()) {
            const Grid &gf = sl->g, &gc = H->lev[l + 1].s->g;
            MgXfer &X = L.X;
            X.N = gf.N; X.sd = gf.sd;
            for (int d = 0; d < 3; ++d) {
                X.ldf[d] = d >= gf.N ? 1 : (d == gf.sd ? gf.lz : gf.pd[d]);
                X.ldc[d] = d >= gc.N ? 1 : (d == gc.sd ? gc.lz : gc.pd[d]);
                X.ncc[d] = gc.nc[d];
            }
            X.P0f = F.P0; X.P0c = H->lev[l + 1].s->F.P0;
        }
        if (getenv("PB200_DEBUG"))
            fprintf(stderr, "[pb200] multigrid level %zu: %d x %d x %d cells, %lld unknowns, lambda_max ~ %.4f\n", l, sl->g.nc[0], sl->g.nc[1], sl->g.nc[2], (long long)sl->dof_bulk, L.lam);
    }
    H->ready = true;
    return PB200_OK;
}

// e = V-cycle(r) on level l (r is left untouched)
static int mg_vcycle(MgHier &H, int l, const FVec &r, const FVec &e)
{
    MgLevel &L = H.lev[l];
    pb200_solver *s = L.s;
    pb200_ctx *ctx = s->ctx;
    FoldSys &F = s->F;
    int rc;
    const int ga = wave_grid(s, kf_mg_add);
    if (l + 1 == (int)H.lev.size()) {
        const double lo = L.lam / H.alpha_c;
        if ((rc = mg_poly(s, r, e, lo, L.lam, 2))) return rc;
        for (int k = 0; k < H.coarse_sweeps; ++k) {
            if ((rc = mg_residual(s, r, e, L.res)) || (rc = mg_poly(s, L.res, L.tmp, lo, L.lam, 2))) return rc;
            kf_mg_add<<<ga, FCH, 0, ctx->stream>>>(F.I, L.tmp, e); LAUNCH_CHECK(ctx);
        }
        return PB200_OK;
    }
    MgLevel &C = H.lev[l + 1];
    pb200_solver *sc = C.s;
    const double lo = L.lam / H.alpha;
    if ((rc = mg_poly(s, r, e, lo, L.lam, H.deg))) return rc;                                  // pre-smoothing from a zero guess
    if ((rc = mg_residual(s, r, e, L.res))) return rc;
    kf_mg_restrict<<<wave_grid(sc, kf_mg_restrict), FCH, 0, ctx->stream>>>(sc->F.I, L.X, F.d.sc[0], sc->F.d.sc[0], L.res.f[0], C.r.f[0]); LAUNCH_CHECK(ctx);
    if ((rc = mg_vcycle(H, l + 1, C.r, C.e))) return rc;
    kf_mg_prolong<<<wave_grid(s, kf_mg_prolong), FCH, 0, ctx->stream>>>(F.I, L.X, F.d.sc[0], sc->F.d.sc[0], C.e.f[0], e.f[0]); LAUNCH_CHECK(ctx);
    if ((rc = mg_residual(s, r, e, L.res)) || (rc = mg_poly(s, L.res, L.tmp, lo, L.lam, H.deg))) return rc;   // post-smoothing
    kf_mg_add<<<ga, FCH, 0, ctx->stream>>>(F.I, L.tmp, e); LAUNCH_CHECK(ctx);
    return PB200_OK;
}

// Multigrid-preconditioned CG on the folded system of s: F.b holds b^, the solution goes to F.x (zero initial guess)
enum { MG_RHO = FS_PAIR0, MG_RR = FS_PAIR0 + 1, MG_RHON = FS_PAIR1 };
static int mg_pcg(pb200_solver *s, const pb200_krylov_opts &o, int *iters, int *conv, double *rnorm_out, double *bnorm_out)
{
    pb200_ctx *ctx = s->ctx;
    FoldSys &F = s->F;
    const Items &I = F.I;
    MgHier &H = *s->mg;
    MgLevel &L0 = H.lev[0];
    double *res = ctx->d_results;
    int rc;
    const int grid = fold_grid(s);
    kf_zero<<<grid, FCH, 0, ctx->stream>>>(I, F.x); LAUNCH_CHECK(ctx);
    kf_resid<<<wave_grid(s, kf_resid), FCH, 0, ctx->stream>>>(I, F.b, F.v, 0, F.r, F.p, F.r0, 0, 1, ctx->d_partials, res + FS_BB, ctx->d_counter); LAUNCH_CHECK(ctx);
    double h[2];
    if ((rc = fetch_results(ctx, FS_BB, 2, h))) return rc;
    const double bnorm = sqrt(h[0]);
    double rnorm = sqrt(h[1]);
    const double tol = fmax(o.rtol * bnorm, o.atol);
    int it = 0, converged = rnorm <= tol ? 1 : 0;
    if (!converged) {
        // one iteration = z = V(r); rho' = (r, z); p = z + (rho' / rho) p; rho = rho'; v = M^ p; alpha = rho / (p, v); x += alpha p; r -= alpha v; (r, r).
        // kf_resid has set p = 0 and rho starts at 1, so the first pass forms p = z without a special case: the body is the same every time and
        // is replayed as ONE CUDA graph from the second iteration on (~100 launches, most of them on tiny coarse levels: pure launch latency).
        const double one = 1.0;
        CUDA_TRY(ctx, cudaMemcpyAsync(res + MG_RHO, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        auto body = [&]() -> int {
            int rc2;
            if ((rc2 = mg_vcycle(H, 0, F.r, L0.e))) return rc2;
            kf_dot<<<wave_grid(s, kf_dot), FCH, 0, ctx->stream>>>(I, F.r, L0.e, ctx->d_partials, res + MG_RHON, ctx->d_counter); LAUNCH_CHECK(ctx);
            kf_mg_p<<<wave_grid(s, kf_mg_p), FCH, 0, ctx->stream>>>(I, res, MG_RHON, MG_RHO, 0, L0.e, F.p); LAUNCH_CHECK(ctx);
            CUDA_TRY(ctx, cudaMemcpyAsync(res + MG_RHO, res + MG_RHON, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            if ((rc2 = fold_apply(s, F.p, F.v, F.v, 1))) return rc2;                         // v = M^ p, (p, v) -> FS_SIG_D
            kf_mg_xr<<<wave_grid(s, kf_mg_xr), FCH, 0, ctx->stream>>>(I, res, MG_RHO, FS_SIG_D, F.p, F.v, F.x, F.r, ctx->d_partials, res + MG_RR, ctx->d_counter); LAUNCH_CHECK(ctx);
            return PB200_OK;
        };
        const bool use_graph = !ctx->profile && !getenv("PB200_NO_GRAPH");
        while (it < o.maxit) {
            if (it == 0 || !use_graph) { if ((rc = body())) return rc; }
            else {
                if (!H.exec) {
                    cudaGraph_t gr = nullptr;
                    const int64_t l0 = ctx->launches, a0 = ctx->apply_launches;
                    CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                    const int rcc = body();
                    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &gr);
                    if (rcc) { if (gr) cudaGraphDestroy(gr); return rcc; }
                    CUDA_TRY(ctx, ce);
                    CUDA_TRY(ctx, cudaGraphInstantiate(&H.exec, gr, 0));
                    cudaGraphDestroy(gr);
                    H.graph_launches = ctx->launches - l0; H.graph_applies = ctx->apply_launches - a0;
                    ctx->launches = l0; ctx->apply_launches = a0;   // capturing launches nothing
                }
                CUDA_TRY(ctx, cudaGraphLaunch(H.exec, ctx->stream));
                ctx->launches += H.graph_launches; ctx->apply_launches += H.graph_applies;
            }
            ++it;
            if ((rc = fetch_results(ctx, MG_RR, 1, h))) return rc;
            rnorm = sqrt(h[0]);
            if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] mg-pcg it %d rnorm %.3e tol %.3e\n", it, rnorm, tol);
            if (rnorm <= tol) { converged = 1; break; }
            if (!(rnorm == rnorm)) break;
        }
    }
    *iters = it; *conv = converged; *rnorm_out = rnorm; *bnorm_out = bnorm;
    return PB200_OK;
}
