#include <algorithm>
// penguin_b200.cu -- C ABI of libpenguin_b200.so (see include/penguin_b200.h for the reference citations).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC penguin_b200.cu -o libpenguin_b200.so
#include "common.cuh"
#include "operators.cuh"
#include "krylov.cuh"
#include "assemble.cuh"
#include "geometry.cuh"
#include "p2p.cuh"
#include "fold.cuh"
#include "fold2.cuh"

#ifndef PB200_POLY_DEFAULT
#define PB200_POLY_DEFAULT 1   // degree of the polynomial preconditioner of the folded CG when PB200_POLY is not set
#endif
#define DISPATCH_N(N_, ...)                  \
    do {                                     \
        if ((N_) == 1) { constexpr int N = 1; __VA_ARGS__; } \
        else if ((N_) == 2) { constexpr int N = 2; __VA_ARGS__; } \
        else { constexpr int N = 3; __VA_ARGS__; } \
    } while (0)

// band kernels: lanes per cell chosen by the size of the band (fold.cuh: band_rows)
#define BAND_POLY_LAUNCH(...) (band_lpc(F.d.nE) == 8 ? kf_band_poly<N, 8><<<band_wgrid(F.d.nE), 256, 0, ctx->stream>>>(__VA_ARGS__) : kf_band_poly<N, 1><<<band_wgrid(F.d.nE), 256, 0, ctx->stream>>>(__VA_ARGS__))
#define BAND_APPLY_LAUNCH(M_, ...) (band_lpc(F.d.nE) == 8 ? kf_apply_band<N, M_, 8><<<band_wgrid(F.d.nE), 256, 0, ctx->stream>>>(__VA_ARGS__) : kf_apply_band<N, M_, 1><<<band_wgrid(F.d.nE), 256, 0, ctx->stream>>>(__VA_ARGS__))

template <typename F> static int team_run(pb200_ctx *tc, F f);   // one host thread per member of a team context (pb200_init_multi, end of this file)
#define IS_TEAM(ctx_) ((ctx_) && (ctx_)->is_team)

// =================================================================================================================
// lifecycle
// =================================================================================================================
static int ctx_common_init(pb200_ctx *c, int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, PB200_ENODEV, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                                  " (libpenguin_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(nullptr, PB200_EINVAL, "bad device index");
    c->device = device;
    CUDA_TRY(c, cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(c, cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_partials, sizeof(double) * RED_MAXK * RED_MAXBLOCKS));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_results, sizeof(double) * RED_SLOTS));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_counter, sizeof(unsigned)));
    CUDA_TRY(c, cudaMemset(c->d_counter, 0, sizeof(unsigned)));
    CUDA_TRY(c, cudaMemset(c->d_results, 0, sizeof(double) * RED_SLOTS));
    CUDA_TRY(c, cudaMallocHost((void **)&c->h_results, sizeof(double) * RED_SLOTS));
    CUDA_TRY(c, cudaEventCreate(&c->ev0));
    CUDA_TRY(c, cudaEventCreate(&c->ev1));
    CUDA_TRY(c, cudaEventCreate(&c->ev2));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_partials2, sizeof(double) * RED_MAXK * RED_MAXBLOCKS));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_counter2, sizeof(unsigned)));
    CUDA_TRY(c, cudaMemset(c->d_counter2, 0, sizeof(unsigned)));
    CUDA_TRY(c, cudaDeviceSynchronize());   // the memsets above ran on the legacy default stream; the library's streams are non-blocking and do not wait for it
    return PB200_OK;
}

extern "C" int pb200_init(pb200_ctx **ctx, int device)
{
    if (!ctx) return set_err(nullptr, PB200_EINVAL, "ctx is NULL");
    pb200_ctx *c = new pb200_ctx();
    int rc = ctx_common_init(c, device);
    if (rc) { delete c; return rc; }
    *ctx = c;
    return PB200_OK;
}

extern "C" int pb200_nccl_unique_id(char id[128])
{
    int rc = nccl_load(nullptr);
    if (rc) return rc;
    pb_ncclUniqueId u;
    NCCL_TRY(nullptr, g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return PB200_OK;
}

extern "C" int pb200_init_dist(pb200_ctx **ctx, int device, int rank, int nranks, const char nccl_id[128])
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return set_err(nullptr, PB200_EINVAL, "bad rank/nranks");
    pb200_ctx *c = new pb200_ctx();
    int rc = ctx_common_init(c, device);
    if (rc) { delete c; return rc; }
    c->rank = rank; c->nranks = nranks;
    if (nranks > 1) {
        rc = nccl_load(c);
        if (rc) { delete c; return rc; }
        pb_ncclUniqueId u;
        memcpy(u.internal, nccl_id, 128);
        int r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
        if (r != 0) { std::string m = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete c; return set_err(nullptr, PB200_ENCCL, m); }
    }
    *ctx = c;
    return PB200_OK;
}

extern "C" int pb200_finalize(pb200_ctx *c)
{
    if (!c) return PB200_OK;
    if (c->is_team) {
        Team *T = c->team;
        team_run(c, [&](int r) { T->ctx[r]->team = nullptr; return pb200_finalize(T->ctx[r]); });   // (teardown of the communicator is collective)
        delete T;
        delete c;
        return PB200_OK;
    }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    p2p_free(c);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_partials); cudaFree(c->d_results); cudaFree(c->d_counter); cudaFreeHost(c->h_results);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    cudaFree(c->d_partials2); cudaFree(c->d_counter2);
    for (cudaEvent_t e : c->pev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
    return PB200_OK;
}
extern "C" const char *pb200_last_error(pb200_ctx *c) { return c ? c->err.c_str() : g_last_error.c_str(); }
extern "C" int pb200_sync(pb200_ctx *c)
{
    if (IS_TEAM(c)) { for (pb200_ctx *m : c->team->ctx) { int rc = pb200_sync(m); if (rc) return rc; } return PB200_OK; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PB200_OK;
}
extern "C" int64_t pb200_launch_count(pb200_ctx *c)
{
    if (IS_TEAM(c)) { int64_t n = 0; for (pb200_ctx *m : c->team->ctx) n += m->launches; return n; }
    return c ? c->launches : 0;
}
extern "C" uint64_t pb200_stream(pb200_ctx *c) { return (uint64_t)(uintptr_t)(IS_TEAM(c) ? c->team->ctx[0]->stream : c->stream); }
extern "C" int pb200_set_profiling(pb200_ctx *c, int enable)
{
    if (!c) return set_err(nullptr, PB200_EINVAL, "NULL ctx");
    if (c->is_team) { for (pb200_ctx *m : c->team->ctx) pb200_set_profiling(m, enable); return PB200_OK; }
    c->profile = enable != 0;
    c->pev_used = 0;
    return PB200_OK;
}


// ---- one process, several GPUs (pb200_init_multi): team versions of the entry points, defined at the end of this file ---------------------
static int team_capacity_create(pb200_ctx *, int, const int *, const double *, const double *, const pb200_levelset *, int, pb200_capacity **);
static int team_capacity_import(pb200_ctx *, int, const int *, const double *, const double *, const double *, const double *, const double *, const double *,
                                const double *, const double *, const double *, const double *, pb200_capacity **);
static int team_capacity_export(pb200_capacity *, double *, double *, double *, double *, double *, double *, double *, double *);
static int team_ops_create(pb200_capacity *, pb200_ops **);
static int team_ops_vec(pb200_ops *, int what, const double *a, const double *b, double *out);
static int team_solver_create(pb200_ctx *, const pb200_solver_desc *, pb200_solver **);
static int team_solver_state(pb200_solver *, const double *in, double *out);
static int team_solver_step(pb200_solver *, const pb200_step_in *, const pb200_krylov_opts *, pb200_step_stats *);
static int team_solver_norms(pb200_solver *, int, const double *, double, int, double *);

// =================================================================================================================
// Capacity
// =================================================================================================================
struct pb200_capacity {
    pb200_ctx *ctx;
    std::vector<pb200_capacity *> parts;   // team handle (pb200_init_multi): one capacity per member rank
    Grid g;
    double *V = nullptr, *Gam = nullptr, *ct = nullptr;
    double *A[PB_MAXD] = {}, *B[PB_MAXD] = {}, *W[PB_MAXD] = {}, *Co[PB_MAXD] = {}, *Cg[PB_MAXD] = {};
    bool has_cg = false;
    // the level set the capacities were built from (pb200_capacity_create): the multigrid preconditioner rebuilds them on coarser meshes (mg.cuh)
    bool has_ls = false;
    int ls_kind = 0, ls_inside = 1, ls_hsdim = 0;
    double ls_hsc = 0.0;
    std::vector<double> ls_c, ls_r;
};

static int cap_alloc(pb200_ctx *ctx, pb200_capacity *c)
{
    int rc;
    const Grid &g = c->g;
    if ((rc = dev_alloc(ctx, &c->V, g.nloc))) return rc;
    if ((rc = dev_alloc(ctx, &c->Gam, g.nloc))) return rc;
    if ((rc = dev_alloc(ctx, &c->ct, g.nloc))) return rc;
    for (int d = 0; d < g.N; ++d) {
        if ((rc = dev_alloc(ctx, &c->A[d], g.nloc))) return rc;
        if ((rc = dev_alloc(ctx, &c->B[d], g.nloc))) return rc;
        if ((rc = dev_alloc(ctx, &c->W[d], g.nloc))) return rc;
        if ((rc = dev_alloc(ctx, &c->Co[d], g.nloc))) return rc;
        // C_gamma only when asked for (compute_centroids / imported): N fields per phase that the solve never reads (6.5 GB per phase at 512^3)
        if (c->has_cg && (rc = dev_alloc(ctx, &c->Cg[d], g.nloc))) return rc;
    }
    return PB200_OK;
}
static int cap_halo(pb200_capacity *c)
{
    std::vector<double *> f = {c->V, c->Gam, c->ct};
    for (int d = 0; d < c->g.N; ++d) { f.push_back(c->A[d]); f.push_back(c->B[d]); f.push_back(c->W[d]); f.push_back(c->Co[d]); }
    return halo_exchange(c->ctx, c->g, f.data(), (int)f.size());
}

extern "C" int pb200_capacity_destroy(pb200_capacity *c);
extern "C" int pb200_capacity_import(pb200_ctx *ctx, int ndim, const int *n, const double *x0, const double *L, const double *V, const double *Gamma,
                                     const double *cell_types, const double *A, const double *B, const double *W, const double *C_omega,
                                     const double *C_gamma, pb200_capacity **out)
{
    if (!ctx || !out) return set_err(ctx, PB200_EINVAL, "NULL argument");
    if (IS_TEAM(ctx)) return team_capacity_import(ctx, ndim, n, x0, L, V, Gamma, cell_types, A, B, W, C_omega, C_gamma, out);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    pb200_capacity *c = new pb200_capacity();
    c->ctx = ctx;
    c->has_cg = C_gamma != nullptr;
    // every failure below leaves through ONE path that releases the partial object (pb200_capacity_destroy frees what exists)
    auto body = [&]() -> int {
        int rc = make_grid(ctx, ndim, n, x0, L, &c->g);
        if (rc || (rc = cap_alloc(ctx, c))) return rc;
        const Grid &g = c->g;
        if ((rc = upload_owned(ctx, g, c->V, V)) || (rc = upload_owned(ctx, g, c->Gam, Gamma)) || (rc = upload_owned(ctx, g, c->ct, cell_types))) return rc;
        for (int d = 0; d < g.N; ++d) {
            if ((rc = upload_owned(ctx, g, c->A[d], A ? A + (int64_t)d * g.nown : nullptr))) return rc;
            if ((rc = upload_owned(ctx, g, c->B[d], B ? B + (int64_t)d * g.nown : nullptr))) return rc;
            if ((rc = upload_owned(ctx, g, c->W[d], W ? W + (int64_t)d * g.nown : nullptr))) return rc;
            if ((rc = upload_owned(ctx, g, c->Co[d], C_omega ? C_omega + (int64_t)d * g.nown : nullptr))) return rc;
            if (c->Cg[d] && (rc = upload_owned(ctx, g, c->Cg[d], C_gamma ? C_gamma + (int64_t)d * g.nown : nullptr))) return rc;
        }
        if ((rc = cap_halo(c))) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return PB200_OK;
    };
    const int rc = body();
    if (rc) { const std::string msg = ctx->err; pb200_capacity_destroy(c); return set_err(ctx, rc, msg); }
    *out = c;
    return PB200_OK;
}

extern "C" int pb200_capacity_create(pb200_ctx *ctx, int ndim, const int *n, const double *x0, const double *L, const pb200_levelset *ls,
                                     int compute_centroids, pb200_capacity **out)
{
    if (!ctx || !out || !ls) return set_err(ctx, PB200_EINVAL, "NULL argument");
    if (IS_TEAM(ctx)) return team_capacity_create(ctx, ndim, n, x0, L, ls, compute_centroids, out);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    pb200_capacity *c = new pb200_capacity();
    c->ctx = ctx;
    c->has_cg = compute_centroids != 0;
    auto body = [&]() -> int {
        int rc = make_grid(ctx, ndim, n, x0, L, &c->g);
        if (rc || (rc = cap_alloc(ctx, c))) return rc;
        GeomOut go;
        go.V = c->V; go.Gam = c->Gam; go.ct = c->ct;
        for (int d = 0; d < PB_MAXD; ++d) { go.A[d] = c->A[d]; go.B[d] = c->B[d]; go.W[d] = c->W[d]; go.Co[d] = c->Co[d]; go.Cg[d] = c->Cg[d]; }
        if ((rc = geometry_build(ctx, c->g, ls, compute_centroids, go))) return rc;
        c->has_ls = true; c->ls_kind = ls->kind; c->ls_inside = ls->fluid_inside; c->ls_hsdim = ls->hs_dim; c->ls_hsc = ls->hs_c;
        if (ls->kind == PB200_LS_BALLS) { c->ls_c.assign(ls->centers, ls->centers + (size_t)ls->nballs * c->g.N); c->ls_r.assign(ls->radii, ls->radii + ls->nballs); }
        if ((rc = cap_halo(c))) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return PB200_OK;
    };
    const int rc = body();
    if (rc) { const std::string msg = ctx->err; pb200_capacity_destroy(c); return set_err(ctx, rc, msg); }
    *out = c;
    return PB200_OK;
}

extern "C" int pb200_capacity_export(pb200_capacity *c, double *V, double *Gamma, double *cell_types, double *A, double *B, double *W, double *C_omega,
                                     double *C_gamma)
{
    if (!c) return set_err(nullptr, PB200_EINVAL, "NULL capacity");
    if (!c->parts.empty()) return team_capacity_export(c, V, Gamma, cell_types, A, B, W, C_omega, C_gamma);
    pb200_ctx *ctx = c->ctx;
    const Grid &g = c->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = download_owned(ctx, g, V, c->V)) || (rc = download_owned(ctx, g, Gamma, c->Gam)) || (rc = download_owned(ctx, g, cell_types, c->ct))) return rc;
    for (int d = 0; d < g.N; ++d) {
        if ((rc = download_owned(ctx, g, A ? A + (int64_t)d * g.nown : nullptr, c->A[d]))) return rc;
        if ((rc = download_owned(ctx, g, B ? B + (int64_t)d * g.nown : nullptr, c->B[d]))) return rc;
        if ((rc = download_owned(ctx, g, W ? W + (int64_t)d * g.nown : nullptr, c->W[d]))) return rc;
        if ((rc = download_owned(ctx, g, C_omega ? C_omega + (int64_t)d * g.nown : nullptr, c->Co[d]))) return rc;
        if (C_gamma && !c->Cg[d]) memset(C_gamma + (int64_t)d * g.nown, 0, sizeof(double) * (size_t)g.nown);   // not computed (compute_centroids = 0)
        else if ((rc = download_owned(ctx, g, C_gamma ? C_gamma + (int64_t)d * g.nown : nullptr, c->Cg[d]))) return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}
extern "C" int pb200_capacity_local(pb200_capacity *c, int *k0, int *k1, int64_t *nloc)
{
    if (!c) return set_err(nullptr, PB200_EINVAL, "NULL capacity");
    if (!c->parts.empty()) { if (k0) *k0 = 0; if (k1) *k1 = c->g.pd[c->g.sd]; if (nloc) *nloc = c->g.ntot; return PB200_OK; }   // team handle: the whole grid
    if (k0) *k0 = c->g.k0;
    if (k1) *k1 = c->g.k1;
    if (nloc) *nloc = c->g.nown;
    return PB200_OK;
}
extern "C" int pb200_capacity_destroy(pb200_capacity *c)
{
    if (!c) return PB200_OK;
    if (!c->parts.empty()) { for (pb200_capacity *p : c->parts) pb200_capacity_destroy(p); delete c; return PB200_OK; }
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    dev_free(c->V); dev_free(c->Gam); dev_free(c->ct);
    for (int d = 0; d < PB_MAXD; ++d) { dev_free(c->A[d]); dev_free(c->B[d]); dev_free(c->W[d]); dev_free(c->Co[d]); dev_free(c->Cg[d]); }
    delete c;
    return PB200_OK;
}

// =================================================================================================================
// DiffusionOps
// =================================================================================================================
struct pb200_ops {
    std::vector<pb200_ops *> parts;        // team handle
    pb200_ctx *ctx;          // (kept here: the destroy path must not reach through `cap`, which a garbage-collected host may have released first)
    pb200_capacity *cap;
    double *Wd[PB_MAXD] = {};
    double *cf = nullptr, *kd = nullptr;   // ConvectionOps (pb200_ops_set_convection): [N][nloc] face flux coefficients, [nloc] interface-velocity diagonal
};
static PhaseDev phase_dev(const pb200_ops *o, const double *Darr, double Dc)
{
    PhaseDev p;
    p.V = o->cap->V; p.Gam = o->cap->Gam;
    for (int d = 0; d < PB_MAXD; ++d) { p.A[d] = o->cap->A[d]; p.B[d] = o->cap->B[d]; p.Wd[d] = o->Wd[d]; }
    p.Darr = Darr; p.Dc = Dc;
    for (int d = 0; d < PB_MAXD; ++d) p.cf[d] = o->cf ? o->cf + (int64_t)d * o->cap->g.nloc : nullptr;
    p.kd = o->kd;
    return p;
}
static inline int sgrid(pb200_ctx *ctx, int64_t n) { return red_grid(ctx, n); }

extern "C" int pb200_ops_destroy(pb200_ops *o);
extern "C" int pb200_ops_create(pb200_capacity *cap, pb200_ops **out)
{
    if (!cap || !out) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!cap->parts.empty()) return team_ops_create(cap, out);
    pb200_ctx *ctx = cap->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    pb200_ops *o = new pb200_ops();
    o->ctx = ctx;
    o->cap = cap;
    for (int d = 0; d < cap->g.N; ++d) {
        int rc = dev_alloc(ctx, &o->Wd[d], cap->g.nloc);
        if (rc) { pb200_ops_destroy(o); return rc; }
        k_wdag<<<sgrid(ctx, cap->g.nloc), RED_THREADS, 0, ctx->stream>>>(cap->g.nloc, cap->W[d], o->Wd[d]);
        LAUNCH_CHECK(ctx);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = o;
    return PB200_OK;
}
extern "C" int pb200_ops_destroy(pb200_ops *o)
{
    if (!o) return PB200_OK;
    if (!o->parts.empty()) { for (pb200_ops *p : o->parts) pb200_ops_destroy(p); delete o; return PB200_OK; }
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    for (int d = 0; d < PB_MAXD; ++d) dev_free(o->Wd[d]);
    dev_free(o->cf); dev_free(o->kd);
    delete o;
    return PB200_OK;
}
// ConvectionOps(capacity, u_omega, u_gamma) (/root/reference/src/operators.jl:194-209): the operators gain the advective terms
//   C_d = D_p diag(S_m A_d u_omega_d) S_m   and   K_d = diag(S_p H' u_gamma)
// as two coefficient sets on the device -- cf_d = S_m (A_d u_omega_d) per direction and kd = 0.5 sum_d S_p^(d) (H' u_gamma) -- that the
// matrix-free rows read (conv_row, operators.cuh).  u_omega: N blocks of n values, u_gamma: N blocks of n values (host, padded grid).
extern "C" int pb200_ops_set_convection(pb200_ops *o, const double *u_omega, const double *u_gamma)
{
    if (!o || !u_omega || !u_gamma) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!o->parts.empty() || o->cap->ctx->nranks > 1) return set_err(o->ctx, PB200_EUNSUPPORTED, "advection-diffusion runs on one GPU so far");
    pb200_ctx *ctx = o->cap->ctx;
    const Grid &g = o->cap->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *uo = nullptr, *ug = nullptr, *zero = nullptr, *q = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &uo, g.nloc * g.N)) || (rc = dev_alloc(ctx, &ug, g.nloc * g.N)) || (rc = dev_alloc(ctx, &zero, g.nloc * g.N)) || (rc = dev_alloc(ctx, &q, g.nloc))) return rc;
    for (int d = 0; d < g.N; ++d) {
        if ((rc = upload_owned(ctx, g, uo + (int64_t)d * g.nloc, u_omega + (int64_t)d * g.nown))) return rc;
        if ((rc = upload_owned(ctx, g, ug + (int64_t)d * g.nloc, u_gamma + (int64_t)d * g.nown))) return rc;
    }
    if (!o->cf && ((rc = dev_alloc(ctx, &o->cf, g.nloc * g.N)) || (rc = dev_alloc(ctx, &o->kd, g.nloc)))) return rc;
    PhaseDev ph = phase_dev(o, nullptr, 1.0);
    DISPATCH_N(g.N, (k_div<N><<<sgrid(ctx, g.nown), RED_THREADS, 0, ctx->stream>>>(g, ph, zero, ug, q)));      // q = H' u_gamma  (-(G' + H') 0 + H' u_gamma)
    LAUNCH_CHECK(ctx);
    DISPATCH_N(g.N, (k_conv_coef<N><<<sgrid(ctx, g.nown), RED_THREADS, 0, ctx->stream>>>(g, ph, uo, q, o->cf, o->kd)));
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(uo); dev_free(ug); dev_free(zero); dev_free(q);
    return PB200_OK;
}
extern "C" int pb200_ops_export_wdag(pb200_ops *o, double *wdag)
{
    if (!o || !wdag) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!o->parts.empty()) return team_ops_vec(o, 0, nullptr, nullptr, wdag);
    pb200_ctx *ctx = o->cap->ctx;
    const Grid &g = o->cap->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    for (int d = 0; d < g.N; ++d) { int rc = download_owned(ctx, g, wdag + (int64_t)d * g.nown, o->Wd[d]); if (rc) return rc; }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

extern "C" int pb200_ops_grad(pb200_ops *o, const double *p, double *out)
{
    if (!o || !p || !out) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!o->parts.empty()) return team_ops_vec(o, 1, p, nullptr, out);
    pb200_ctx *ctx = o->cap->ctx;
    const Grid &g = o->cap->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *u = nullptr, *gm = nullptr, *q = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &u, g.nloc)) || (rc = dev_alloc(ctx, &gm, g.nloc)) || (rc = dev_alloc(ctx, &q, g.nloc * g.N))) return rc;
    if ((rc = upload_owned(ctx, g, u, p)) || (rc = upload_owned(ctx, g, gm, p + g.nown))) return rc;
    double *fl[2] = {u, gm};
    if ((rc = halo_exchange(ctx, g, fl, 2))) return rc;
    PhaseDev ph = phase_dev(o, nullptr, 1.0);
    dim3 grid(sgrid(ctx, g.nown), g.N);
    DISPATCH_N(g.N, (k_grad<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, ph, u, gm, q)));
    LAUNCH_CHECK(ctx);
    for (int d = 0; d < g.N; ++d)
        if ((rc = download_owned(ctx, g, out + (int64_t)d * g.nown, q + (int64_t)d * g.nloc))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(u); dev_free(gm); dev_free(q);
    return PB200_OK;
}
extern "C" int pb200_ops_div(pb200_ops *o, const double *qo, const double *qg, double *out)
{
    if (!o || !qo || !qg || !out) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!o->parts.empty()) return team_ops_vec(o, 2, qo, qg, out);
    pb200_ctx *ctx = o->cap->ctx;
    const Grid &g = o->cap->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *a = nullptr, *b = nullptr, *r = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &a, g.nloc * g.N)) || (rc = dev_alloc(ctx, &b, g.nloc * g.N)) || (rc = dev_alloc(ctx, &r, g.nloc))) return rc;
    std::vector<double *> fl;
    for (int d = 0; d < g.N; ++d) {
        if ((rc = upload_owned(ctx, g, a + (int64_t)d * g.nloc, qo + (int64_t)d * g.nown))) return rc;
        if ((rc = upload_owned(ctx, g, b + (int64_t)d * g.nloc, qg + (int64_t)d * g.nown))) return rc;
        fl.push_back(a + (int64_t)d * g.nloc); fl.push_back(b + (int64_t)d * g.nloc);
    }
    if ((rc = halo_exchange(ctx, g, fl.data(), (int)fl.size()))) return rc;
    PhaseDev ph = phase_dev(o, nullptr, 1.0);
    DISPATCH_N(g.N, (k_div<N><<<sgrid(ctx, g.nown), RED_THREADS, 0, ctx->stream>>>(g, ph, a, b, r)));
    LAUNCH_CHECK(ctx);
    if ((rc = download_owned(ctx, g, out, r))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    dev_free(a); dev_free(b); dev_free(r);
    return PB200_OK;
}

/* the two coefficient sets of pb200_ops_set_convection back on the host: cf (ndim * n: S_m A_d u_omega_d) and kd (n: the diagonal of 0.5 sum_d K_d) */
extern "C" int pb200_ops_export_convection(pb200_ops *o, double *cf, double *kd)
{
    if (!o) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!o->parts.empty() || !o->cf) return set_err(o->ctx, PB200_EINVAL, "no convection on this operator handle");
    pb200_ctx *ctx = o->cap->ctx;
    const Grid &g = o->cap->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    for (int d = 0; d < g.N && cf; ++d)
        if ((rc = download_owned(ctx, g, cf + (int64_t)d * g.nown, o->cf + (int64_t)d * g.nloc))) return rc;
    if (kd && (rc = download_owned(ctx, g, kd, o->kd))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

// =================================================================================================================
// Solver
// =================================================================================================================
struct MgHier;
struct pb200_solver;
static void mg_free(pb200_solver *s);
struct pb200_solver {
    pb200_ctx *ctx;
    MgHier *mg = nullptr;                  // multigrid hierarchy of the folded system (mg.cuh), built on first use
    std::vector<pb200_solver *> parts;     // team handle
    int team_nblk = 0;                     // blocks of the state vector (2 mono, 4 diph)
    Grid g;
    SysParams sp;
    pb200_ops *o1 = nullptr, *o2 = nullptr;
    PhaseDev p1, p2;
    double *D1arr = nullptr, *D2arr = nullptr;
    BorderDev bd;
    double *bvals[6] = {};
    bool masks_dirty = true;
    bool values_dirty = false;           // border values changed, kinds did not: refresh ufix only
    bool have_generic = false;           // Krylov vectors of the generic path allocated
    unsigned char *m1 = nullptr, *m2 = nullptr;
    double *ufix1 = nullptr, *ufix2 = nullptr;
    double *Tw[2] = {}, *Tg[2] = {};
    // older states T^(n-1), T^(n-2), ... for the extrapolated initial guess; the buffers rotate with Tw / Tg (no copies)
    std::vector<double *> histW[2], histG[2];
    int n_prev = 0;                      // how many of them are valid
    double prev_dt = 0.0;
    int nf = 1;
    MVec x, b, r, r0, p, ph, v, s, sh, t, dinv;
    double *gK = nullptr;
    double *fS[2][2] = {}, *gS[2] = {};
    int64_t dof_bulk = 0, dof_ifc = 0;
    ApplyCoef diag_key = {-1, -1, -1, -1, -1};
    std::vector<double *> owned;
    FoldSys F;
    BandHalo bh;
    long long *Kcell = nullptr;   // sorted list of the rows whose right-hand side has a known part (built with the masks)
    int nK = 0;
    cudaStream_t copy_stream = nullptr;      // pb200_solver_get_state_async
    cudaEvent_t state_ready = nullptr, copy_done = nullptr;
    bool copy_pending = false;
};

static int solver_vec(pb200_solver *s, MVec *v)
{
    for (int f = 0; f < KV_MAXF; ++f) v->f[f] = nullptr;
    for (int f = 0; f < s->nf; ++f) {
        int rc = dev_alloc(s->ctx, &v->f[f], s->g.nloc);
        if (rc) return rc;
        s->owned.push_back(v->f[f]);
    }
    return PB200_OK;
}

extern "C" int pb200_solver_destroy(pb200_solver *s);
extern "C" int pb200_solver_create(pb200_ctx *ctx, const pb200_solver_desc *d, pb200_solver **out)
{
    if (!ctx || !d || !out || !d->ops1) return set_err(ctx, PB200_EINVAL, "NULL argument");
    if (d->phase_type == PB200_DIPH && !d->ops2) return set_err(ctx, PB200_EINVAL, "diphasic solver needs ops2");
    if (IS_TEAM(ctx)) return team_solver_create(ctx, d, out);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    pb200_solver *s = new pb200_solver();
    s->ctx = ctx;
    s->g = d->ops1->cap->g;
    const Grid &g = s->g;
    auto body = [&]() -> int {   // every failure leaves through pb200_solver_destroy on the partial object (below)
    s->o1 = d->ops1; s->o2 = d->ops2;
    s->sp.phase_type = d->phase_type; s->sp.time_type = d->time_type; s->sp.ifc_kind = d->ifc_kind;
    s->sp.alpha = d->alpha; s->sp.beta = d->beta;
    s->sp.a1 = d->alpha1; s->sp.a2 = d->alpha2; s->sp.b1 = d->beta1; s->sp.b2 = d->beta2;
    if (d->phase_type == PB200_MONO) {
        if (d->ifc_kind == PB200_BC_DIRICHLET) { s->sp.alpha = 1.0; s->sp.beta = 0.0; }
        else if (d->ifc_kind == PB200_BC_NEUMANN) { s->sp.alpha = 0.0; s->sp.beta = 1.0; }
        else if (d->ifc_kind != PB200_BC_ROBIN) return set_err(ctx, PB200_EINVAL, "mono interface condition must be Dirichlet, Neumann or Robin");
    } else {
        if (d->alpha1 == 0.0) return set_err(ctx, PB200_EUNSUPPORTED, "ScalarJump with alpha1 == 0 is not supported");
        const Grid &g2 = d->ops2->cap->g;
        if (g2.N != g.N || g2.ntot != g.ntot) return set_err(ctx, PB200_EINVAL, "phase capacities must share the same mesh");
    }
    int rc;
    if (d->D1_arr) { if ((rc = dev_alloc(ctx, &s->D1arr, g.nloc)) || (rc = upload_owned(ctx, g, s->D1arr, d->D1_arr))) return rc; }
    if (d->D2_arr) { if ((rc = dev_alloc(ctx, &s->D2arr, g.nloc)) || (rc = upload_owned(ctx, g, s->D2arr, d->D2_arr))) return rc; }
    {
        double *fl[2] = {s->D1arr, s->D2arr};
        if ((rc = halo_exchange(ctx, g, fl, 2))) return rc;
    }
    s->p1 = phase_dev(d->ops1, s->D1arr, d->D1);
    if (d->ops2) s->p2 = phase_dev(d->ops2, s->D2arr, d->D2); else s->p2 = s->p1;
    for (int k = 0; k < 6; ++k) { s->bd.kind[k] = PB200_BC_NONE; s->bd.present[k] = 0; s->bd.value[k] = 0.0; s->bd.values[k] = nullptr; }
    const bool diph = d->phase_type == PB200_DIPH;
    s->nf = diph ? 3 : (s->sp.beta != 0.0 ? 2 : 1);
    CUDA_TRY(ctx, cudaMalloc((void **)&s->m1, (size_t)g.nloc));
    CUDA_TRY(ctx, cudaMemsetAsync(s->m1, 0, (size_t)g.nloc, ctx->stream));
    CUDA_TRY(ctx, cudaMalloc((void **)&s->m2, (size_t)g.nloc));
    CUDA_TRY(ctx, cudaMemsetAsync(s->m2, 0, (size_t)g.nloc, ctx->stream));
    if ((rc = dev_alloc(ctx, &s->ufix1, g.nloc)) || (rc = dev_alloc(ctx, &s->ufix2, g.nloc))) return rc;
    for (int ph = 0; ph < (diph ? 2 : 1); ++ph)
        if ((rc = dev_alloc(ctx, &s->Tw[ph], g.nloc)) || (rc = dev_alloc(ctx, &s->Tg[ph], g.nloc))) return rc;
    if ((rc = dev_alloc(ctx, &s->gK, g.nloc))) return rc;
    // x, b: both solve paths.  The Krylov vectors of the generic path (9 more: 30 GB at 512^3 diphasic) are allocated when that path first runs.
    MVec *vs[] = {&s->x, &s->b};
    for (MVec *v : vs) if ((rc = solver_vec(s, v))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
    };
    const int rcb = body();
    if (rcb) { const std::string msg = ctx->err; pb200_solver_destroy(s); return set_err(ctx, rcb, msg); }
    *out = s;
    return PB200_OK;
}

extern "C" int pb200_solver_destroy(pb200_solver *s)
{
    if (!s) return PB200_OK;
    if (!s->parts.empty()) {
        pb200_ctx *tc = s->ctx;
        team_run(tc, [&](int r) { return pb200_solver_destroy(s->parts[r]); });   // (in parallel: destroying may synchronise streams that wait on peers)
        delete s;
        return PB200_OK;
    }
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    mg_free(s);
    for (double *p : s->owned) cudaFree(p);
    fold_free(s->F);
    if (s->Kcell) cudaFree(s->Kcell);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); cudaEventDestroy(s->state_ready); cudaEventDestroy(s->copy_done); }
    dev_free(s->D1arr); dev_free(s->D2arr); dev_free(s->ufix1); dev_free(s->ufix2); dev_free(s->gK);
    for (int k = 0; k < 6; ++k) dev_free(s->bvals[k]);
    for (int a = 0; a < 2; ++a) { for (double *p : s->histW[a]) cudaFree(p); for (double *p : s->histG[a]) cudaFree(p); }
    for (int a = 0; a < 2; ++a) { dev_free(s->Tw[a]); dev_free(s->Tg[a]); dev_free(s->gS[a]); for (int b = 0; b < 2; ++b) dev_free(s->fS[a][b]); }
    cudaFree(s->m1); cudaFree(s->m2);
    delete s;
    return PB200_OK;
}

static int64_t side_cells(const Grid &g, int side)
{
    const int dim = (side == PB200_LEFT || side == PB200_RIGHT) ? 1 : (side == PB200_BOTTOM || side == PB200_TOP) ? 0 : 2;
    int64_t n = 1;
    for (int d = 0; d < g.N; ++d) if (d != dim) n *= g.nc[d];
    return n;
}

extern "C" int pb200_solver_set_border(pb200_solver *s, int side, int kind, double value, const double *values)
{
    if (!s || side < 0 || side > 5) return set_err(nullptr, PB200_EINVAL, "bad side");
    if (!s->parts.empty()) {   // the side array is indexed by GLOBAL cell coordinates on every rank: the same host array goes to every member
        for (pb200_solver *p : s->parts) { int rc = pb200_solver_set_border(p, side, kind, value, values); if (rc) return set_err(s->ctx, rc, g_last_error); }
        return PB200_OK;
    }
    pb200_ctx *ctx = s->ctx;
    const int dim = (side == PB200_LEFT || side == PB200_RIGHT) ? 1 : (side == PB200_BOTTOM || side == PB200_TOP) ? 0 : 2;
    if (dim >= s->g.N) return PB200_OK;   // keys of absent dimensions never match a cell (src/solver.jl:379-409)
    if (kind == PB200_BC_NEUMANN && s->g.N == 1 && ctx->nranks > 1)
        return set_err(ctx, PB200_EUNSUPPORTED, "1-D Neumann border rows are supported on one rank only");
    // Neumann (>= 2-D) and Robin borders are no-ops in the reference (src/solver.jl:471-498): record as NONE -- but the key is PRESENT,
    // which is what a Periodic row on the opposite side tests (src/solver.jl:458).  In 1-D the Neumann row is real (src/solver.jl:471-493).
    if (kind != PB200_BC_DIRICHLET && kind != PB200_BC_PERIODIC && !(kind == PB200_BC_NEUMANN && s->g.N == 1)) kind = PB200_BC_NONE;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // The reference's time loop calls BC_border_mono! every step (src/solver/diffusion.jl:291-293).  Masks, folded system and captured
    // graphs depend on WHICH rows are pinned (kind / presence per side), not on the pinned values: a call that only changes values
    // refreshes ufix with one light kernel at the next step; a call that changes nothing costs nothing.
    const bool same_kind = s->bd.present[side] == 1 && s->bd.kind[side] == kind;
    const bool had_arr = s->bd.values[side] != nullptr, has_arr = values && kind == PB200_BC_DIRICHLET;
    if (!same_kind) s->masks_dirty = true;
    else if (has_arr || had_arr || s->bd.value[side] != value) s->values_dirty = true;
    s->bd.present[side] = 1;
    s->bd.kind[side] = kind;
    s->bd.value[side] = value;
    if (has_arr) {
        int64_t n = side_cells(s->g, side);
        if (!s->bvals[side]) CUDA_TRY(ctx, cudaMalloc((void **)&s->bvals[side], sizeof(double) * (size_t)n));
        CUDA_TRY(ctx, cudaMemcpyAsync(s->bvals[side], values, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        s->bd.values[side] = s->bvals[side];
    } else s->bd.values[side] = nullptr;
    return PB200_OK;
}

extern "C" int pb200_solver_set_state(pb200_solver *s, const double *x)
{
    if (!s || !x) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return team_solver_state(s, x, nullptr);
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int np = s->sp.phase_type == PB200_DIPH ? 2 : 1;
    for (int ph = 0; ph < np; ++ph) {
        if ((rc = upload_owned(ctx, g, s->Tw[ph], x + (int64_t)(2 * ph) * g.nown))) return rc;
        if ((rc = upload_owned(ctx, g, s->Tg[ph], x + (int64_t)(2 * ph + 1) * g.nown))) return rc;
    }
    s->n_prev = 0;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}
extern "C" int pb200_solver_get_state(pb200_solver *s, double *x)
{
    if (!s || !x) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return team_solver_state(s, nullptr, x);
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int np = s->sp.phase_type == PB200_DIPH ? 2 : 1;
    for (int ph = 0; ph < np; ++ph) {
        if ((rc = download_owned(ctx, g, x + (int64_t)(2 * ph) * g.nown, s->Tw[ph]))) return rc;
        if ((rc = download_owned(ctx, g, x + (int64_t)(2 * ph + 1) * g.nown, s->Tg[ph]))) return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PB200_OK;
}

// check_convergence (src/convergence.jl:59-93) on the device state: out[0..3] = norms over all fluid (full + cut), full, cut, empty cells
extern "C" int pb200_solver_error_norms(pb200_solver *s, int phase, const double *u_ana, double p, int relative, double out[4])
{
    if (!s || !u_ana || !out) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return team_solver_norms(s, phase, u_ana, p, relative, out);
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    const int np = s->sp.phase_type == PB200_DIPH ? 2 : 1;
    if (phase < 0 || phase >= np) return set_err(ctx, PB200_EINVAL, "phase index out of range");
    const bool is_inf = std::isinf(p);
    if (!is_inf && !(p > 0.0)) return set_err(ctx, PB200_EINVAL, "the norm order p must be positive");
    if (is_inf && ctx->nranks > 1) return set_err(ctx, PB200_EUNSUPPORTED, "L-infinity error norms are rank-local: not available on a slab-partitioned grid");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const pb200_capacity *cap = (phase == 0 ? s->o1 : s->o2)->cap;
    double *ua = nullptr;
    unsigned long long *mx = nullptr;
    int rc;
    if ((rc = dev_alloc(ctx, &ua, g.nloc))) return rc;
    if ((rc = upload_owned(ctx, g, ua, u_ana))) return rc;
    CUDA_TRY(ctx, cudaMalloc((void **)&mx, sizeof(unsigned long long) * 6));
    CUDA_TRY(ctx, cudaMemsetAsync(mx, 0, sizeof(unsigned long long) * 6, ctx->stream));
    k_err_norms<<<sgrid(ctx, g.nown), RED_THREADS, 0, ctx->stream>>>(g, cap->ct, cap->V, ua, s->Tw[phase], p, relative ? 1 : 0, is_inf ? 1 : 0, mx, ctx->d_partials,
                                                                    ctx->d_results + SL_TMP, ctx->d_counter);
    LAUNCH_CHECK(ctx);
    if ((rc = allreduce_results(ctx, SL_TMP, 4))) return rc;
    double sums[4], hm[6];
    if ((rc = fetch_results(ctx, SL_TMP, 4, sums))) return rc;
    CUDA_TRY(ctx, cudaMemcpy(hm, mx, sizeof(hm), cudaMemcpyDeviceToHost));   // (bit patterns of doubles)
    cudaFree(mx);
    dev_free(ua);
    auto cls = [&](int a, int b) -> double {   // class a (and b): the reference's lp_norm / relative_lp_norm
        if (!is_inf) { const double S = sums[a] + (b >= 0 ? sums[b] : 0.0); return pow(S / sums[3], 1.0 / p); }
        const double me = b >= 0 ? fmax(hm[a], hm[b]) : hm[a], mu = b >= 0 ? fmax(hm[3 + a], hm[3 + b]) : hm[3 + a];
        if (!relative) return me;
        // errors[idx] / u_ana[idx] on two Julia VECTORS is err * pinv(u_ana) (a matrix): max |err_i u_j| / sum u^2 (src/convergence.jl:19)
        return me * mu / (sums[a] + (b >= 0 ? sums[b] : 0.0));
    };
    out[0] = cls(0, 1); out[1] = cls(0, -1); out[2] = cls(1, -1); out[3] = cls(2, -1);
    return PB200_OK;
}

extern "C" int pb200_solver_get_state_async(pb200_solver *s, double *x)
{
    if (!s || !x) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return team_solver_state(s, nullptr, x);   // team handle: the slabs are reassembled on the host, synchronously
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!s->copy_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&s->state_ready, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaEventRecord(s->state_ready, ctx->stream));          // the state of the last step is complete here
    CUDA_TRY(ctx, cudaStreamWaitEvent(s->copy_stream, s->state_ready, 0));
    const int np = s->sp.phase_type == PB200_DIPH ? 2 : 1;
    for (int ph = 0; ph < np; ++ph) {
        CUDA_TRY(ctx, cudaMemcpyAsync(x + (int64_t)(2 * ph) * g.nown, s->Tw[ph] + g.plane, sizeof(double) * (size_t)g.nown, cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(x + (int64_t)(2 * ph + 1) * g.nown, s->Tg[ph] + g.plane, sizeof(double) * (size_t)g.nown, cudaMemcpyDeviceToHost, s->copy_stream));
    }
    CUDA_TRY(ctx, cudaEventRecord(s->copy_done, s->copy_stream));
    s->copy_pending = true;
    return PB200_OK;
}
extern "C" int pb200_solver_wait_state(pb200_solver *s)
{
    if (!s) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return PB200_OK;
    if (s->copy_pending) { CUDA_TRY(s->ctx, cudaEventSynchronize(s->copy_done)); s->copy_pending = false; }
    return PB200_OK;
}

static bool has_slave_rows(const pb200_solver *s)
{
    return s->g.N == 1 && (s->bd.kind[PB200_BOTTOM] == PB200_BC_NEUMANN || s->bd.kind[PB200_TOP] == PB200_BC_NEUMANN);
}

// ---- operator application on Krylov vectors ------------------------------------------------------------------------
static int apply_op(pb200_solver *s, const ApplyCoef &ac, const MVec &in, const MVec &out)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    int rc;
    if ((rc = halo_exchange(ctx, g, in.f, s->nf))) return rc;
    const int grid = sgrid(ctx, g.nown);
    const bool slaves = has_slave_rows(s);
    if (slaves) {
        k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m1, in.f[0], 0); LAUNCH_CHECK(ctx);
        if (s->sp.phase_type == PB200_DIPH) { k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m2, in.f[1], 0); LAUNCH_CHECK(ctx); }
    }
    prof_mark(ctx, PB_PROF_APPLY);
    ctx->apply_launches++;
    if (s->sp.phase_type == PB200_MONO) {
        GamSpec gs = {s->nf == 2 ? in.f[1] : nullptr, 1.0, nullptr, 0.0, 0.0};
        DISPATCH_N(g.N, (k_apply_mono<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->sp, ac, s->m1, in.f[0], gs, out.f[0],
                                                                                 s->nf == 2 ? out.f[1] : nullptr)));
    } else {
        GamSpec g1 = {in.f[2], s->sp.a2 / s->sp.a1, nullptr, 0.0, 0.0};
        GamSpec g2 = {in.f[2], 1.0, nullptr, 0.0, 0.0};
        DISPATCH_N(g.N, (k_apply_diph<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->p2, s->sp, ac, s->m1, s->m2, in.f[0], g1, in.f[1], g2,
                                                                                 out.f[0], out.f[1], out.f[2])));
    }
    LAUNCH_CHECK(ctx);
    prof_mark(ctx, PB_PROF_APPLY);
    if (slaves) {   // Krylov vectors carry zeros on non-free entries
        k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m1, in.f[0], 1); LAUNCH_CHECK(ctx);
        if (s->sp.phase_type == PB200_DIPH) { k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m2, in.f[1], 1); LAUNCH_CHECK(ctx); }
    }
    return PB200_OK;
}


// =================================================================================================================
// folded fast path (fold.cuh): build + Krylov
// =================================================================================================================
static bool fold_eligible(const pb200_solver *s)
{
    const SysParams &sp = s->sp;
    if (s->p1.kd || s->p2.kd) return false;   // advection: the system is not symmetric -- reference rows, BiCGSTAB
    if (has_slave_rows(s)) return false;   // 1-D Neumann border rows: eliminated unknowns that follow a neighbour -- generic path only
    if (sp.phase_type == PB200_MONO) {
        if (sp.beta == 0.0) return true;                       // Dirichlet interface: T_gamma known, SPD bulk system
        return sp.beta > 0.0 && sp.alpha >= 0.0;               // Robin / Neumann rows symmetrise with positive factors
    }
    return sp.a1 != 0.0 && sp.a2 / sp.a1 > 0.0 && sp.b1 > 0.0 && sp.b2 > 0.0;
}

static int fold_alloc_vec(pb200_solver *s, FVec *v)
{
    FoldSys &F = s->F;
    int rc;
    for (int f = 0; f < 3; ++f) v->f[f] = nullptr;
    for (int f = 0; f < F.d.nbulk; ++f) if ((rc = dev_alloc(s->ctx, &v->f[f], F.nlocq))) return rc;   // re-pitched layout (fold.cuh: FoldDev)
    if (F.d.has_w && (rc = dev_alloc(s->ctx, &v->f[2], F.d.nB > 0 ? F.d.nB : 1))) return rc;
    return PB200_OK;
}

// two-pass compaction + sort of a marked cell list
template <typename Launch>
static int fold_compact(pb200_ctx *ctx, Launch launch, long long **list, int *n)
{
    int *d_cnt = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&d_cnt, sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(int), ctx->stream));
    launch((long long *)nullptr, d_cnt, 0);
    LAUNCH_CHECK(ctx);
    int cnt = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    *list = nullptr;
    if (cnt > 0) {
        CUDA_TRY(ctx, cudaMalloc((void **)list, sizeof(long long) * (size_t)cnt));
        CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(int), ctx->stream));
        launch(*list, d_cnt, cnt);
        LAUNCH_CHECK(ctx);
        thrust::sort(thrust::cuda::par.on(ctx->stream), *list, *list + cnt);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_cnt);
    return PB200_OK;
}


static inline int band_grid(int n);
static inline int band_wgrid(int n);
static inline int band_lpc(int n);
// extreme eigenvalues of the band block M^_BB by power iteration (set-up, O(band) work) -> coefficients of the band preconditioner
static int fold_band_spectrum(pb200_solver *s)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    FoldSys &F = s->F;
    const FoldDev &d = F.d;
    int rc;
    const size_t nB = d.nB > 0 ? d.nB : 1;
    CUDA_TRY(ctx, cudaMalloc((void **)&F.dz, sizeof(double) * 3 * nB));
    CUDA_TRY(ctx, cudaMemsetAsync(F.dz, 0, sizeof(double) * 3 * nB, ctx->stream));
    const int gb = band_wgrid(d.nE), gB = band_grid(d.nB);
    const StopCrit none = {0.0, 0.0, -1};
    double *res = ctx->d_results;
    // the iteration vector lives in F.v (zero outside the band, restored to zero afterwards)
    auto power = [&](double shift, double sign, int iters, double *lam) -> int {
        kf_band_seed<<<gB, 128, 0, ctx->stream>>>(d, F.dz); LAUNCH_CHECK(ctx);
        double nrm2 = 0.0, h[1];
        // normalise the seed
        kf_band_put<<<gB, 128, 0, ctx->stream>>>(d, F.v, F.dz, 1.0, 0, res, none); LAUNCH_CHECK(ctx);
        double rq = 0.0;
        for (int it = 0; it < iters; ++it) {
            // ||x||^2 via (x, 1 x + 0 A x)
            DISPATCH_N(g.N, (BAND_POLY_LAUNCH(g, d, F.v, F.dz, 1.0, 0.0, ctx->d_partials, res + SL_TMP, ctx->d_counter, res, none)));
            LAUNCH_CHECK(ctx);
            if ((rc = allreduce_results(ctx, SL_TMP, 1))) return rc;
            if ((rc = fetch_results(ctx, SL_TMP, 1, h))) return rc;
            nrm2 = h[0];
            if (!(nrm2 > 0.0)) { *lam = 0.0; return PB200_OK; }
            kf_band_put<<<gB, 128, 0, ctx->stream>>>(d, F.v, F.dz, 1.0 / sqrt(nrm2), 0, res, none); LAUNCH_CHECK(ctx);
            // y = (shift + sign A) x ; Rayleigh quotient (x, y)
            DISPATCH_N(g.N, (BAND_POLY_LAUNCH(g, d, F.v, F.dz, shift, sign, ctx->d_partials, res + SL_TMP, ctx->d_counter, res, none)));
            LAUNCH_CHECK(ctx);
            if ((rc = allreduce_results(ctx, SL_TMP, 1))) return rc;
            if ((rc = fetch_results(ctx, SL_TMP, 1, h))) return rc;
            rq = h[0];
            kf_band_put<<<gB, 128, 0, ctx->stream>>>(d, F.v, F.dz, 1.0, 0, res, none); LAUNCH_CHECK(ctx);
        }
        *lam = rq;
        return PB200_OK;
    };
    double lmax = 0.0, lshift = 0.0;
    if ((rc = power(0.0, 1.0, 40, &lmax))) return rc;
    const double hi = 1.05 * lmax;
    if ((rc = power(hi, -1.0, 60, &lshift))) return rc;
    double lmin = hi - lshift;
    // clean the work vector
    CUDA_TRY(ctx, cudaMemsetAsync(F.dz, 0, sizeof(double) * 3 * nB, ctx->stream));
    kf_band_put<<<gB, 128, 0, ctx->stream>>>(d, F.v, F.dz, 0.0, 0, res, none); LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    F.band_lmin = lmin; F.band_lmax = lmax;
    if (!(lmax > 0.0) || !(lmin > 0.0) || !(lmin < lmax)) { F.prec = false; return PB200_OK; }
    const double lo = 0.7 * lmin;
    // two Chebyshev steps on [lo, hi]: q(t) = a0 + a1 t, positive for t < lo + hi
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    const double rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
    F.pa0 = (1.0 + rho1 * rho0) / theta + 2.0 * rho1 / delta;
    F.pa1 = -2.0 * rho1 / (delta * theta);
    F.prec = true;
    if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] band block spectrum ~ [%.4f, %.4f], prec q(t) = %.4f %+.4f t\n", lmin, lmax, F.pa0, F.pa1);
    return PB200_OK;
}

typedef CUresult (*pb_encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pb_encode_tiled_t pb_encode_tiled();
static int fold_tmap(pb200_solver *s, const double *ptr, CUtensorMap *out, int kind = 0);
static int fold_build(pb200_solver *s, const ApplyCoef &ac)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    FoldSys &F = s->F;
    mg_free(s);          // (the coarse levels mirror this system)
    fold_free(F);
    FoldDev &d = F.d;
    memset(&d, 0, sizeof(d));
    const SysParams &sp = s->sp;
    const bool diph = sp.phase_type == PB200_DIPH;
    d.N = g.N; d.nbulk = diph ? 2 : 1; d.has_w = diph || sp.beta != 0.0;
    d.c = ac.c; d.cVc = ac.cV / ac.c;
    if (diph) {
        const double kap = sp.a2 / sp.a1;
        d.kap[0] = kap; d.kap[1] = 1.0; d.s[0] = sp.b1 / kap; d.s[1] = sp.b2; d.mwc = 0.0; d.wrow = 1.0;
        d.m[0] = s->m1; d.m[1] = s->m2; d.mw = s->m2;
    } else {
        d.kap[0] = d.kap[1] = 1.0; d.s[0] = d.s[1] = 1.0;
        d.mwc = d.has_w ? sp.alpha / sp.beta : 0.0;
        d.wrow = d.has_w ? 1.0 / (ac.c2 * sp.beta) : 1.0;
        d.m[0] = s->m1; d.m[1] = s->m1; d.mw = s->m1;
    }
    d.ph[0] = s->p1; d.ph[1] = s->p2;
    {
        // layout of the Krylov vectors: x rows padded to a multiple of 32 doubles (256-byte aligned tile rows, TMA-describable)
        const long long ld0 = g.sd == 0 ? g.lz : g.pd[0], ld1 = g.N < 2 ? 1 : (g.sd == 1 ? g.lz : g.pd[1]), ld2 = g.N < 3 ? 1 : g.lz;
        long long P0 = ld0;
        if (g.N >= 2 && !getenv("PB200_NO_REPITCH")) P0 = (ld0 + 31) / 32 * 32;
        d.ld0 = ld0; d.dP = P0 - ld0;
        d.sq[0] = 1; d.sq[1] = P0; d.sq[2] = P0 * ld1;
        F.P0 = P0; F.nlocq = P0 * ld1 * ld2;
        F.planeq = g.N == 1 ? 1 : (g.N == 2 ? P0 : P0 * ld1);
    }
    int rc;
    {
        unsigned char *ml[2] = {s->m1, s->m2};
        if ((rc = halo_exchange_bytes(ctx, g, ml, 2))) return rc;
    }
    for (int p = 0; p < d.nbulk; ++p) {
        if ((rc = dev_alloc(ctx, &F.sc[p], g.nloc))) return rc;
        d.sc[p] = F.sc[p];
        for (int dd = 0; dd < g.N; ++dd) { if ((rc = dev_alloc(ctx, &F.off[p][dd], g.nloc))) return rc; d.off[p][dd] = F.off[p][dd]; }
    }
    const int grid = red_grid(ctx, g.nloc);
    std::vector<long long> hB;
    if (d.has_w) {
        const unsigned char *mw = d.mw;
        const long long nloc = g.nloc;
        cudaStream_t st = ctx->stream;
        if ((rc = fold_compact(ctx, [&](long long *list, int *cnt, int cap) { kf_mark_band<<<grid, RED_THREADS, 0, st>>>(nloc, mw, list, cnt, cap); }, &F.Bcell, &d.nB))) return rc;
        d.Bcell = F.Bcell;
        hB.resize(d.nB);
        if (d.nB) CUDA_TRY(ctx, cudaMemcpy(hB.data(), F.Bcell, sizeof(long long) * (size_t)d.nB, cudaMemcpyDeviceToHost));
        int lo = 0, hi = 0;
        for (long long l : hB) { if (l < g.plane) ++lo; else if (l >= g.plane + g.nown) ++hi; }
        d.nBlo = lo; d.nBown = d.nB - lo - hi;
        CUDA_TRY(ctx, cudaMalloc((void **)&F.bord, sizeof(int) * (size_t)g.nloc));
        CUDA_TRY(ctx, cudaMemsetAsync(F.bord, 0xFF, sizeof(int) * (size_t)g.nloc, ctx->stream));
        if (d.nB) { kf_bord_fill<<<(d.nB + 255) / 256, 256, 0, ctx->stream>>>(d.nB, F.Bcell, F.bord); LAUNCH_CHECK(ctx); }
        d.bord = F.bord;
        CUDA_TRY(ctx, cudaMalloc((void **)&F.Linv, sizeof(double) * 5 * (size_t)(d.nB > 0 ? d.nB : 1)));
        CUDA_TRY(ctx, cudaMemsetAsync(F.Linv, 0, sizeof(double) * 5 * (size_t)(d.nB > 0 ? d.nB : 1), ctx->stream));
        d.Linv = F.Linv;
        fold_band_ranges(ctx, g, hB, d.nBlo, d.nBown, &s->bh);
    }
    {
        const int nBhi = d.nB - d.nBlo - d.nBown;
        int mb = d.nBlo > nBhi ? d.nBlo : nBhi;
        if (s->bh.lo_sendn > mb) mb = s->bh.lo_sendn;
        if (s->bh.hi_sendn > mb) mb = s->bh.hi_sendn;
        if ((rc = p2p_setup(ctx, (size_t)d.nbulk * F.planeq + mb))) return rc;   // collective (every rank builds its folded system here)
    }
    const int gown = red_grid(ctx, g.nown);
    DISPATCH_N(g.N, (kf_diag<N><<<gown, RED_THREADS, 0, ctx->stream>>>(g, d)));
    LAUNCH_CHECK(ctx);
    if ((rc = halo_exchange(ctx, g, F.sc, d.nbulk))) return rc;
    if (d.has_w && (rc = fold_band_halo(ctx, F, s->bh, F.Linv, 5))) return rc;
    if (d.has_w) {
        const int *bord = F.bord;
        cudaStream_t st = ctx->stream;
        Grid gg = g;
        if (g.N == 1) { if ((rc = fold_compact(ctx, [&](long long *list, int *cnt, int cap) { kf_mark_E<1><<<gown, RED_THREADS, 0, st>>>(gg, bord, list, cnt, cap); }, &F.Ecell, &d.nE))) return rc; }
        else if (g.N == 2) { if ((rc = fold_compact(ctx, [&](long long *list, int *cnt, int cap) { kf_mark_E<2><<<gown, RED_THREADS, 0, st>>>(gg, bord, list, cnt, cap); }, &F.Ecell, &d.nE))) return rc; }
        else { if ((rc = fold_compact(ctx, [&](long long *list, int *cnt, int cap) { kf_mark_E<3><<<gown, RED_THREADS, 0, st>>>(gg, bord, list, cnt, cap); }, &F.Ecell, &d.nE))) return rc; }
        d.Ecell = F.Ecell;
        d.nEp = (d.nE + 31) / 32 * 32;
        const size_t nE = d.nEp > 0 ? d.nEp : 32;
        CUDA_TRY(ctx, cudaMalloc((void **)&F.EB, sizeof(int) * nE));
        CUDA_TRY(ctx, cudaMalloc((void **)&F.EnbrB, sizeof(int) * 2 * g.N * nE));
        CUDA_TRY(ctx, cudaMalloc((void **)&F.Eblk, sizeof(double) * (1 + 2 * g.N) * 9 * nE));
        CUDA_TRY(ctx, cudaMalloc((void **)&F.Efix, nE));
        d.EB = F.EB; d.EnbrB = F.EnbrB; d.Eblk = F.Eblk; d.Efix = F.Efix;
        // (small band) band heads of the fused iteration (fold2.cuh): cell -> E index, band cell -> E index, band couplings of the bulk rows.
        // Several ranks: implemented (collective choice below) but OFF unless PB200_BANDFUSE_MULTI is set -- on 2 GPUs the graph-replayed iteration of the
        // weak-scaled 2-D bench ran 2.3x SLOWER with the heads (335 vs 146 us) although every kernel timed alone was as fast or faster; not understood yet.
        if (band_lpc(d.nE) == 8 && (ctx->nranks == 1 || getenv("PB200_BANDFUSE_MULTI"))) {
            CUDA_TRY(ctx, cudaMalloc((void **)&F.eord, sizeof(int) * (size_t)g.nloc));
            CUDA_TRY(ctx, cudaMemsetAsync(F.eord, 0xFF, sizeof(int) * (size_t)g.nloc, ctx->stream));
            CUDA_TRY(ctx, cudaMalloc((void **)&F.EofB, sizeof(int) * (size_t)(d.nB > 0 ? d.nB : 1)));
            CUDA_TRY(ctx, cudaMemsetAsync(F.EofB, 0xFF, sizeof(int) * (size_t)(d.nB > 0 ? d.nB : 1), ctx->stream));   // (band cells in ghost planes have no E row)
            CUDA_TRY(ctx, cudaMalloc((void **)&F.EnbrE, sizeof(int) * 2 * g.N * nE));
            CUDA_TRY(ctx, cudaMalloc((void **)&F.ya, sizeof(double) * 2 * nE));
            CUDA_TRY(ctx, cudaMemsetAsync(F.ya, 0, sizeof(double) * 2 * nE, ctx->stream));
            if (d.nE) { kf_eord_fill<<<(d.nE + 255) / 256, 256, 0, ctx->stream>>>(d.nE, F.Ecell, F.bord, F.eord, F.EofB); LAUNCH_CHECK(ctx); }
            d.eord = F.eord; d.EofB = F.EofB; d.ya = F.ya; d.EnbrE = F.EnbrE;
        }
        if (d.nEp == 0) d.nEp = 32;
        {   // local planes of the slab dimension whose tiles are interior class (fused kernel): see kf_tile_records (ghost flag)
            const int Tsd = g.N == 1 ? FTILE : (g.N == 2 ? 32 : 4);
            d.fix_lo = ctx->rank > 0 ? Tsd : 0;
            d.fix_hi = ctx->rank < ctx->nranks - 1 ? ((g.lz - 2) / Tsd) * Tsd : g.lz;
        }
    }
    DISPATCH_N(g.N, (kf_off<N><<<gown, RED_THREADS, 0, ctx->stream>>>(g, d)));
    LAUNCH_CHECK(ctx);
    {
        double *fl[2 * PB_MAXD];
        int nf = 0;
        for (int p = 0; p < d.nbulk; ++p) for (int dd = 0; dd < g.N; ++dd) fl[nf++] = F.off[p][dd];
        if ((rc = halo_exchange(ctx, g, fl, nf))) return rc;
    }
    if (d.has_w && d.nE > 0) {
        DISPATCH_N(g.N, (kf_blocks<N><<<(d.nE + 127) / 128, 128, 0, ctx->stream>>>(g, d)));
        LAUNCH_CHECK(ctx);
        if (F.EnbrE) { kf_enbre_fill<<<(d.nE + 255) / 256, 256, 0, ctx->stream>>>(g, d.nE, d.nEp, F.Ecell, F.EnbrB, F.eord, F.EnbrE); LAUNCH_CHECK(ctx); }
        if (F.EnbrE && d.nB > 0 && band_lpc(d.nE) == 8) {   // small band: band-indexed copies for the update head
            d.nBp = (d.nB + 31) / 32 * 32;
            CUDA_TRY(ctx, cudaMalloc((void **)&F.Bq, sizeof(long long) * (size_t)d.nBp));
            CUDA_TRY(ctx, cudaMalloc((void **)&F.Bidx, sizeof(int) * (size_t)(1 + 4 * g.N) * d.nBp));
            CUDA_TRY(ctx, cudaMalloc((void **)&F.Bblk, sizeof(double) * (size_t)(1 + 2 * g.N) * 9 * d.nBp));
            kf_band_index<<<(d.nB + 127) / 128, 128, 0, ctx->stream>>>(g, d, F.Bq, F.Bidx, F.Bblk, d.nBp); LAUNCH_CHECK(ctx);
            d.Bq = F.Bq; d.Bidx = F.Bidx; d.Bblk = F.Bblk;
        }
    }
    // Band heads (fold2.cuh) are a COLLECTIVE choice: every rank must hold a small band (or none) whose cells AND their face neighbours all lie in
    // tiles of the fused kernel (interior class: Efix has every bit set) -- then no band row needs a ghost plane and the heads replace the band launches
    // on every rank (bench.py's weak-scaled configs[1]: one interface per slab, away from the slab faces).
    F.bandfuse_ok = false;
    if (d.has_w) {
        double bad = (d.nE > 0 && F.eord == nullptr) ? 1.0 : 0.0;
        if (d.nE > 0 && bad == 0.0 && ctx->nranks > 1) {
            std::vector<unsigned char> hfix((size_t)d.nE);
            CUDA_TRY(ctx, cudaMemcpyAsync(hfix.data(), F.Efix, (size_t)d.nE, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            const unsigned full = (1u << (1 + 2 * g.N)) - 1u;
            for (unsigned char v : hfix) if ((v & full) != full) { bad = 1.0; break; }
        }
        if (d.nE > 0 && d.nB > 0 && F.Bq == nullptr) bad = 1.0;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_results + SL_TMP, &bad, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = allreduce_results(ctx, SL_TMP, 1))) return rc;
        if ((rc = fetch_results(ctx, SL_TMP, 1, &bad))) return rc;
        F.bandfuse_ok = bad == 0.0;
        if (getenv("PB200_REPORT_HEADS") && ctx->rank == 0) fprintf(stderr, "[pb200] band heads: %s (%d rank%s, %d band / fringe rows on rank 0)\n", F.bandfuse_ok ? "on" : "off", ctx->nranks, ctx->nranks > 1 ? "s" : "", d.nE);
    }
    // active tile list + per-tile coefficient census
    {
        Items &I = F.I;
        memset(&I, 0, sizeof(I));
        I.sd = g.sd; I.lz = g.lz;
        I.glo = ctx->rank > 0 ? 1 : 0; I.ghi = ctx->rank < ctx->nranks - 1 ? 1 : 0;
        I.ld0 = g.sd == 0 ? g.lz : g.pd[0];
        I.ld1 = g.N < 2 ? 1 : (g.sd == 1 ? g.lz : g.pd[1]);
        I.ld2 = g.N < 3 ? 1 : g.lz;
        if (g.N == 1) { I.T0 = FTILE; I.T1 = 1; I.T2 = 1; I.shx = 8; I.kx = FCH; I.ky = 0; I.kz = 0; I.tym = 1; I.ustride = FCH; }
        else if (g.N == 2) { I.T0 = 32; I.T1 = 32; I.T2 = 1; I.shx = 5; I.kx = 0; I.ky = 1; I.kz = 0; I.tym = FU; I.ustride = I.ld0; }
        else { I.T0 = 32; I.T1 = 8; I.T2 = 4; I.shx = 5; I.kx = 0; I.ky = 0; I.kz = 1; I.tym = 1; I.ustride = I.ld0 * I.ld1; }
        I.P0 = F.P0; I.nq = F.nlocq; I.nl = g.nloc;
        I.ustrideq = g.N == 1 ? FCH : (g.N == 2 ? F.P0 : F.P0 * I.ld1);
        I.nt0 = (int)((I.ld0 + I.T0 - 1) / I.T0); I.nt1 = (int)((I.ld1 + I.T1 - 1) / I.T1);
        const long long nt2 = (I.ld2 + I.T2 - 1) / I.T2;
        const long long ntile = (long long)I.nt0 * I.nt1 * nt2;
        if (ntile >= (1ll << 30)) return set_err(ctx, PB200_EUNSUPPORTED, "more than 2^30 tiles per rank");
        const int wchunks = d.has_w ? (d.nBown + FTILE - 1) / FTILE : 0;
        const long long tot = (long long)d.nbulk * ntile + wchunks;
        int *flags = nullptr, *d_cnt = nullptr;
        CUDA_TRY(ctx, cudaMalloc((void **)&flags, sizeof(int) * (size_t)(d.nbulk * ntile)));
        CUDA_TRY(ctx, cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)(d.nbulk * ntile), ctx->stream));
        CUDA_TRY(ctx, cudaMalloc((void **)&d_cnt, sizeof(int)));
        CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(int), ctx->stream));
        CUDA_TRY(ctx, cudaMalloc((void **)&F.items, sizeof(int) * (size_t)(tot > 0 ? tot : 1)));
        kf_tile_flags<<<gown, RED_THREADS, 0, ctx->stream>>>(g, I, d.nbulk, d.m[0], d.m[1], ntile, flags);
        LAUNCH_CHECK(ctx);
        kf_tile_list<<<red_grid(ctx, tot), RED_THREADS, 0, ctx->stream>>>(d.nbulk, ntile, flags, wchunks, F.items, d_cnt);
        LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(&F.nitems, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (F.nitems > 0) {
            unsigned *it = (unsigned *)F.items;
            thrust::sort(thrust::cuda::par.on(ctx->stream), it, it + F.nitems);
            ctx->launches++;
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(flags); cudaFree(d_cnt);
        I.it = F.items; I.n = F.nitems; I.wlo = d.nBlo; I.whi = d.nBlo + d.nBown;
        const size_t ni = F.nitems > 0 ? F.nitems : 1;
        CUDA_TRY(ctx, cudaMalloc((void **)&F.rec, sizeof(TileRec) * ni));
        if (F.nitems > 0) { kf_tile_records<<<(F.nitems + 255) / 256, 256, 0, ctx->stream>>>(I, F.rec); LAUNCH_CHECK(ctx); }
        I.rec = F.rec;
        CUDA_TRY(ctx, cudaMalloc((void **)&F.uni, ni));
        CUDA_TRY(ctx, cudaMalloc((void **)&F.ucoef, sizeof(double) * PB_MAXD * ni));
        I.uni = F.uni; I.ucoef = F.ucoef;
        int gm = F.nitems; if (gm > ctx->sm_count * 8) gm = ctx->sm_count * 8; if (gm < 1) gm = 1;
        DISPATCH_N(g.N, (kf_tile_meta<N><<<gm, FCH, 0, ctx->stream>>>(g, d, I, F.uni, F.ucoef, getenv("PB200_EXACT_TILES") ? 0.0 : 1e-12, ctx->d_partials, ctx->d_results + SL_TMP, ctx->d_counter)));
        LAUNCH_CHECK(ctx);
        double cnt[3];
        if ((rc = fetch_results(ctx, SL_TMP, 3, cnt))) return rc;
        F.cells_uniform = (long long)(cnt[0] + 0.5); F.cells_general = (long long)(cnt[1] + 0.5); F.cells_fast = (long long)(cnt[2] + 0.5);
        // Static load balance of the operator apply.  The item-loop kernels hand item i to block i mod grid.  For the apply kernel a
        // tile with streamed coefficients or partial validity costs several times a constant-coefficient interior tile, and where those
        // sit in index order decides how many of them one block draws (measured: 296 -> 248 us per apply at 384^3).  The apply therefore
        // walks a SECOND list: slow tiles first, then the fast ones, each class in index order -- every block gets the same number of
        // tiles of each class +-1 for any grid size, and the order is fixed, so its reduction stays deterministic.  The streaming vector
        // kernels keep the index-ordered list: all tiles cost them the same, and neighbouring blocks on neighbouring tiles share DRAM
        // pages (the class order cost them 25 % at 384^3).
        F.IA = I; F.IAg = I; F.IG1 = I; F.IAf = I; F.IAgen = I; F.IFall = I; F.IGall = I; F.IAi_all = I;
        F.IG1nw = I; F.IG1nw.n = 0;
        F.IAg.n = 0; F.IG1.n = 0; F.IAf.n = 0; F.IAgen.n = 0; F.IFall.n = 0; F.IGall.n = 0; F.IAi_all.n = 0;
        if (F.nitems > 0) {
            const int n = F.nitems;
            std::vector<int> hi(n);
            std::vector<unsigned char> hu(n);
            std::vector<TileRec> hr(n);
            CUDA_TRY(ctx, cudaMemcpy(hi.data(), F.items, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
            CUDA_TRY(ctx, cudaMemcpy(hu.data(), F.uni, (size_t)n, cudaMemcpyDeviceToHost));
            CUDA_TRY(ctx, cudaMemcpy(hr.data(), F.rec, sizeof(TileRec) * (size_t)n, cudaMemcpyDeviceToHost));
            // a sub-list of the items with its own records / flags / constants
            auto make_list = [&](const std::vector<int> &ids, Items &out) -> int {
                out = I;
                out.n = (int)ids.size();
                if (ids.empty()) return PB200_OK;
                const size_t na = ids.size();
                int *d_it = nullptr; TileRec *d_rec = nullptr; unsigned char *d_uni = nullptr; double *d_uc = nullptr;
                CUDA_TRY(ctx, cudaMalloc((void **)&d_it, sizeof(int) * na)); F.list_mem.push_back(d_it);
                CUDA_TRY(ctx, cudaMalloc((void **)&d_rec, sizeof(TileRec) * na)); F.list_mem.push_back(d_rec);
                CUDA_TRY(ctx, cudaMalloc((void **)&d_uni, na)); F.list_mem.push_back(d_uni);
                CUDA_TRY(ctx, cudaMalloc((void **)&d_uc, sizeof(double) * PB_MAXD * na)); F.list_mem.push_back(d_uc);
                // Stream-ordered: a plain cudaMemcpy from pageable memory returns when the data is STAGED, the DMA to the device may still be
                // running, and the kernels below are launched on a non-blocking stream that does not wait for the legacy default stream
                // (seen on B200: a run-dependent suffix of a list read as zeros = tile 0 -> stray work on tile 0, garbage on the real tiles).
                CUDA_TRY(ctx, cudaMemcpyAsync(d_it, ids.data(), sizeof(int) * na, cudaMemcpyHostToDevice, ctx->stream));
                out.it = d_it;
                kf_tile_records<<<((int)na + 255) / 256, 256, 0, ctx->stream>>>(out, d_rec); LAUNCH_CHECK(ctx);
                out.rec = d_rec; out.uni = d_uni; out.ucoef = d_uc;
                int ga = (int)na; if (ga > ctx->sm_count * 8) ga = ctx->sm_count * 8;
                DISPATCH_N(g.N, (kf_tile_meta<N><<<ga, FCH, 0, ctx->stream>>>(g, d, out, d_uni, d_uc, getenv("PB200_EXACT_TILES") ? 0.0 : 1e-12, ctx->d_partials, ctx->d_results + SL_TMP, ctx->d_counter)));
                LAUNCH_CHECK(ctx);
                double cnt2[3];
                return fetch_results(ctx, SL_TMP, 3, cnt2);
            };
            // cost-class order of the bulk tiles (slow first): all / interior class / ghost class (box reaches a neighbour rank's ghost plane)
            const bool reorder = !getenv("PB200_NO_REORDER");
            std::vector<int> all, ghost, g1, fast_i, gen_i, fast_all, gen_all, inner_all;
            // 3-D: the apply's lists walk the tiles BRICK by brick (8 x 8 x 8 tiles = 256 x 64 x 32 cells) instead of in index order, so that the tiles in
            // flight at one time form a compact block whose halo planes are shared while they are in L2.  Measured with ncu at 1024 x 1024 x 128 (DRAM
            // reads of one fused apply, 4.1 GB algorithmic): index order, 2 blocks per SM x 2 stages 6.8 GB; index order, 1 block per SM x 4 stages 5.5 GB;
            // bricks, 1 block per SM 4.9 GB (tools/r2_ncu_*.sh).  PB200_ORDER=runs: consecutive z / y neighbours handed to ONE block in runs (Items.run,
            // f3_item) -- measured worse (7.4 GB), kept as an experiment.  The order is fixed, so the reductions stay deterministic.
            std::vector<int> perm(n);
            for (int i = 0; i < n; ++i) perm[i] = i;
            int run3 = 1;
            const char *order = getenv("PB200_ORDER");
            if (g.N == 3 && !getenv("PB200_NO_BRICKS") && !(order && !strcmp(order, "index"))) {
                std::vector<long long> key(n);
                if (order && !strcmp(order, "runs")) {
                    int ry = 2, rz = 8;
                    if (const char *e = getenv("PB200_RUN")) sscanf(e, "%d,%d", &ry, &rz);
                    if (ry < 1) ry = 1;
                    if (rz < 1) rz = 1;
                    run3 = ry * rz;
                    const long long nt2l = (I.ld2 + I.T2 - 1) / I.T2;
                    for (int i = 0; i < n; ++i) {
                        if (hr[i].f >= 2) { key[i] = (1ll << 62) + i; continue; }
                        const long long t0 = hr[i].ox / I.T0, t1 = hr[i].oy / I.T1, t2 = hr[i].oz / I.T2;
                        key[i] = ((long long)hr[i].f << 58) + ((((t0 * ((I.nt1 + ry - 1) / ry) + t1 / ry) * ((nt2l + rz - 1) / rz) + t2 / rz) * rz + t2 % rz) * ry + t1 % ry);
                    }
                } else {
                    int bd[3] = {8, 8, 8};
                    if (const char *e = getenv("PB200_BRICK")) sscanf(e, "%d,%d,%d", &bd[0], &bd[1], &bd[2]);
                    for (int q = 0; q < 3; ++q) if (bd[q] < 1) bd[q] = 1;
                    const long long nb0 = (I.nt0 + bd[0] - 1) / bd[0], nb1 = (I.nt1 + bd[1] - 1) / bd[1];
                    for (int i = 0; i < n; ++i) {
                        if (hr[i].f >= 2) { key[i] = (1ll << 62) + i; continue; }
                        const long long t0 = hr[i].ox / I.T0, t1 = hr[i].oy / I.T1, t2 = hr[i].oz / I.T2;
                        const long long brick = (t0 / bd[0]) + nb0 * ((t1 / bd[1]) + nb1 * (t2 / bd[2]));
                        const long long inb = (t0 % bd[0]) + bd[0] * ((t1 % bd[1]) + bd[1] * (t2 % bd[2]));
                        key[i] = ((long long)hr[i].f << 58) + brick * (long long)(bd[0] * bd[1] * bd[2]) + inb;
                    }
                }
                std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
            }
            for (int cls = 0; cls < 2; ++cls)
                for (int ii = 0; ii < n; ++ii) {
                    const int i = perm[ii];
                    if (hr[i].f >= 2) continue;     // (w chunks are not the dense apply's business)
                    // every cell valid, one constant per direction, no band cell: the pipelined kernel (kf3_apply).  (A tile in which a field lives on
                    // cut cells only has all-zero coefficients -- "constant" -- and still holds band cells, whose z carries the band correction.)
                    const bool fast = (hu[i] & 1) && hr[i].full && !(hu[i] & 2);
                    const int c = (!reorder || fast) ? 1 : 0;
                    if (c != cls) continue;
                    all.push_back(hi[i]);
                    (fast ? fast_all : gen_all).push_back(hi[i]);
                    if (hr[i].ghost) ghost.push_back(hi[i]);
                    else { (fast ? fast_i : gen_i).push_back(hi[i]); inner_all.push_back(hi[i]); }
                }
            // pointwise p / x update of the fused iteration: ghost-class tiles + the compact interface unknowns (index order)
            for (int i = 0; i < n; ++i) if (hr[i].f >= 2 || hr[i].ghost) g1.push_back(hi[i]);
            std::vector<int> g1nw;                                                                    // (band heads: the heads update the interface unknowns)
            for (int i = 0; i < n; ++i) if (hr[i].f < 2 && hr[i].ghost) g1nw.push_back(hi[i]);
            if ((rc = make_list(g1nw, F.IG1nw))) return rc;
            if ((rc = make_list(all, F.IA)) || (rc = make_list(ghost, F.IAg)) || (rc = make_list(g1, F.IG1)) || (rc = make_list(fast_i, F.IAf)) ||
                (rc = make_list(gen_i, F.IAgen)) || (rc = make_list(fast_all, F.IFall)) || (rc = make_list(gen_all, F.IGall)) || (rc = make_list(inner_all, F.IAi_all))) return rc;
            F.IA.run = F.IAg.run = F.IAf.run = F.IAgen.run = F.IFall.run = F.IGall.run = F.IAi_all.run = run3;
            if (getenv("PB200_DBG_LISTS")) {   // debugging: every sub-list must carry the base list's records / flags / constants for its items
                std::vector<double> hc((size_t)n * PB_MAXD);
                CUDA_TRY(ctx, cudaMemcpy(hc.data(), F.ucoef, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
                std::map<int, int> pos;
                for (int i = 0; i < n; ++i) pos[hi[i]] = i;
                const Items *Ls[8] = {&F.IA, &F.IAg, &F.IG1, &F.IAf, &F.IAgen, &F.IFall, &F.IGall, &F.IAi_all};
                const char *nm[8] = {"IA", "IAg", "IG1", "IAf", "IAgen", "IFall", "IGall", "IAi_all"};
                for (int q = 0; q < 8; ++q) {
                    const Items &Lq = *Ls[q];
                    if (Lq.n == 0) continue;
                    std::vector<int> li(Lq.n); std::vector<TileRec> lr(Lq.n); std::vector<unsigned char> lu(Lq.n); std::vector<double> lc((size_t)Lq.n * PB_MAXD);
                    CUDA_TRY(ctx, cudaMemcpy(li.data(), Lq.it, sizeof(int) * (size_t)Lq.n, cudaMemcpyDeviceToHost));
                    CUDA_TRY(ctx, cudaMemcpy(lr.data(), Lq.rec, sizeof(TileRec) * (size_t)Lq.n, cudaMemcpyDeviceToHost));
                    CUDA_TRY(ctx, cudaMemcpy(lu.data(), Lq.uni, (size_t)Lq.n, cudaMemcpyDeviceToHost));
                    CUDA_TRY(ctx, cudaMemcpy(lc.data(), Lq.ucoef, sizeof(double) * lc.size(), cudaMemcpyDeviceToHost));
                    long bad_rec = 0, bad_uni = 0, bad_c = 0; int first = -1;
                    for (int j = 0; j < Lq.n; ++j) {
                        const int b = pos[li[j]];
                        const bool r_ok = lr[j].base == hr[b].base && lr[j].baseq == hr[b].baseq && lr[j].f == hr[b].f && lr[j].full == hr[b].full && lr[j].ox == hr[b].ox && lr[j].oz == hr[b].oz;
                        const bool bulk = hr[b].f < 2;
                        const bool u_ok = !bulk || lu[j] == hu[b];
                        bool c_ok = true;
                        if (bulk) for (int dd = 0; dd < g.N; ++dd) c_ok = c_ok && lc[(size_t)j * PB_MAXD + dd] == hc[(size_t)b * PB_MAXD + dd];
                        if (!r_ok) ++bad_rec; if (!u_ok) ++bad_uni; if (!c_ok) ++bad_c;
                        if ((!r_ok || !u_ok || !c_ok) && first < 0) first = j;
                    }
                    fprintf(stderr, "[pb200] list %s: %d items, bad records %ld, bad flags %ld, bad constants %ld (first bad item %d)\n", nm[q], Lq.n, bad_rec, bad_uni, bad_c, first);
                }
            }
        }
    }
    FVec *vs[] = {&F.x, &F.b, &F.r, &F.p, &F.v};
    for (FVec *v : vs) if ((rc = fold_alloc_vec(s, v))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    F.key[0] = ac.cV; F.key[1] = ac.c; F.key[2] = ac.c2;
    F.built = true;
    // TMA staging needs global strides that are multiples of 16 bytes (an even pitch: the re-pitched layout) and >= 2-D grids
    F.tma_ok = g.N >= 2 && (F.P0 % 2 == 0) && !getenv("PB200_NO_TMA") && pb_encode_tiled() != nullptr;
    if (F.tma_ok) { CUtensorMap probe; if (fold_tmap(s, F.x.f[0], &probe)) { F.tma_ok = false; cudaGetLastError(); } }
    F.pipe = F.tma_ok && !getenv("PB200_NO_PIPE");
    // the band preconditioner is a COLLECTIVE decision (its set-up and every iteration contain reductions over the ranks): it is used
    // when ANY rank holds band cells, and ranks without band cells simply contribute zeros
    if (d.has_w && !getenv("PB200_NO_BAND_PREC")) {
        double ne = (double)d.nE;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_results + SL_TMP, &ne, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = allreduce_results(ctx, SL_TMP, 1))) return rc;
        if ((rc = fetch_results(ctx, SL_TMP, 1, &ne))) return rc;
        if (ne > 0.5 && (rc = fold_band_spectrum(s))) return rc;
    }
    return PB200_OK;
}

static inline int fold_grid(pb200_solver *s) { int b = s->F.nitems; int cap = s->ctx->sm_count * 8; if (b > cap) b = cap; if (b < 1) b = 1; return b; }
// Grid of a kernel that loops over the work items: EXACTLY one wave (resident blocks per SM x SMs).  With a larger grid the second,
// partial wave leaves SMs idle at the tail (ncu: 1.6 waves, SM-active 74 % of elapsed for the 48-register apply kernel at 2048^2).
template <typename K>
static inline int wave_grid(pb200_solver *s, K kernel)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, FCH, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 4; }
    int b = s->F.nitems, cap = s->ctx->sm_count * nb;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return b;
}
static inline int band_grid(int n) { int b = (n + 127) / 128; if (b > RED_MAXBLOCKS) b = RED_MAXBLOCKS; if (b < 1) b = 1; return b; }
// kernels that put one warp on one band cell: 1024-thread blocks, so that every cell is in flight at once while the number of
// blocks (= serialised ticket atomics and partial sums of the fused reduction) stays small
// one thread per band / fringe cell; small bands (2-D) use 64-thread blocks so that their few thousand cells still spread over every SM
static inline int band_lpc(int n) { return n < 148 * 512 ? 8 : 1; }
static inline int band_wgrid(int n) { const long long t = (long long)n * band_lpc(n); long long b = (t + 255) / 256; if (b > 2048) b = 2048; if (b < 1) b = 1; return (int)b; }

// ghost planes of the bulk fields and ghost entries of the compact w of one Krylov vector, ONE NCCL group (one launch)
static int fold_halo(pb200_solver *s, const FVec &x, cudaStream_t st = nullptr)
{
    pb200_ctx *ctx = s->ctx;
    if (ctx->nranks == 1) return PB200_OK;
    if (!st) st = ctx->stream;
    prof_mark(ctx, PB_PROF_XCHG);
    struct Mark { pb200_ctx *c; ~Mark() { prof_mark(c, PB_PROF_XCHG); } } mark_{ctx};
    const Grid &g = s->g;
    const FoldSys &F = s->F;
    const int nB = F.d.nB, nBlo = F.d.nBlo, nBown = F.d.nBown, nBhi = nB - nBlo - nBown;
    const BandHalo &bh = s->bh;
    const size_t cnt = (size_t)F.planeq;   // Krylov vectors: planes of the re-pitched layout
    if (ctx->p2p && ctx->p2p->on) {
        // peer-memory path (p2p.cuh): one kernel stores the boundary data into the neighbours' mailboxes and unpacks what they stored here
        P2PState *P = ctx->p2p;
        HaloDirArgs lo, hi;
        memset(&lo, 0, sizeof(lo)); memset(&hi, 0, sizeof(hi));
        const bool band = F.d.has_w && nB > 0;
        size_t total = 0;
        if (ctx->rank > 0) {
            lo.active = 1; lo.remote = P->peer[ctx->rank - 1]; lo.remote_dir = 1; lo.local_dir = 0;
            for (int f = 0; f < F.d.nbulk; ++f) { lo.send[lo.nseg] = {x.f[f] + F.planeq, nullptr, (int)cnt}; lo.recv[lo.nseg] = {nullptr, x.f[f], (int)cnt}; ++lo.nseg; total += cnt; }
            if (band) { lo.send[lo.nseg] = {x.f[2] + bh.lo_send0, nullptr, bh.lo_sendn}; lo.recv[lo.nseg] = {nullptr, x.f[2], nBlo}; ++lo.nseg; }
        }
        if (ctx->rank < ctx->nranks - 1) {
            hi.active = 1; hi.remote = P->peer[ctx->rank + 1]; hi.remote_dir = 0; hi.local_dir = 1;
            for (int f = 0; f < F.d.nbulk; ++f) {
                hi.send[hi.nseg] = {x.f[f] + (long long)(g.lz - 2) * F.planeq, nullptr, (int)cnt};
                hi.recv[hi.nseg] = {nullptr, x.f[f] + (long long)(g.lz - 1) * F.planeq, (int)cnt};
                ++hi.nseg; total += cnt;
            }
            if (band) { hi.send[hi.nseg] = {x.f[2] + bh.hi_send0, nullptr, bh.hi_sendn}; hi.recv[hi.nseg] = {nullptr, x.f[2] + nBlo + nBown, nBhi}; ++hi.nseg; }
        }
        int blocks = (int)(total / 8192); if (blocks < 1) blocks = 1; if (blocks > 64) blocks = 64;
        k_p2p_halo<<<blocks, 512, 0, st>>>(P->mbox, P->zone_doubles, lo, hi);
        LAUNCH_CHECK(ctx);
        return PB200_OK;
    }
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int f = 0; f < F.d.nbulk; ++f) {
        double *p = x.f[f];
        if (ctx->rank > 0) {
            NCCL_TRY(ctx, g_nccl.Send(p + F.planeq, cnt, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, st));
            NCCL_TRY(ctx, g_nccl.Recv(p, cnt, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, st));
        }
        if (ctx->rank < ctx->nranks - 1) {
            NCCL_TRY(ctx, g_nccl.Send(p + (long long)(g.lz - 2) * F.planeq, cnt, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, st));
            NCCL_TRY(ctx, g_nccl.Recv(p + (long long)(g.lz - 1) * F.planeq, cnt, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, st));
        }
    }
    if (F.d.has_w && nB > 0) {
        double *p = x.f[2];
        if (ctx->rank > 0) {
            if (bh.lo_sendn) NCCL_TRY(ctx, g_nccl.Send(p + bh.lo_send0, (size_t)bh.lo_sendn, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, st));
            if (nBlo) NCCL_TRY(ctx, g_nccl.Recv(p, (size_t)nBlo, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, st));
        }
        if (ctx->rank < ctx->nranks - 1) {
            if (bh.hi_sendn) NCCL_TRY(ctx, g_nccl.Send(p + bh.hi_send0, (size_t)bh.hi_sendn, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, st));
            if (nBhi) NCCL_TRY(ctx, g_nccl.Recv(p + nBlo + nBown, (size_t)nBhi, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, st));
        }
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    return PB200_OK;
}

// ---- TMA descriptors of the Krylov vectors (fold2.cuh) ---------------------------------------------------------------------------
static pb_encode_tiled_t pb_encode_tiled()
{
    static pb_encode_tiled_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (pb_encode_tiled_t)p;
        else cudaGetLastError();
    }
    return fn;
}
// descriptor of one bulk field of a Krylov vector: an N-d tensor of doubles (P0, ld1[, ld2]) read in boxes of tile + halo
static int fold_tmap(pb200_solver *s, const double *ptr, CUtensorMap *out, int kind)   // kind 0: tile + halo box, 1: tile
{
    FoldSys &F = s->F;
    auto it = F.tmaps.find(std::make_pair(ptr, kind));
    if (it != F.tmaps.end()) { *out = it->second; return PB200_OK; }
    pb_encode_tiled_t enc = pb_encode_tiled();
    if (!enc) return set_err(s->ctx, PB200_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available");
    const int N = s->g.N;
    const cuuint64_t dims[3] = {(cuuint64_t)F.P0, (cuuint64_t)F.I.ld1, (cuuint64_t)F.I.ld2};
    const cuuint64_t strides[2] = {(cuuint64_t)F.P0 * 8, (cuuint64_t)F.P0 * (cuuint64_t)F.I.ld1 * 8};
    cuuint32_t box2[2] = {(cuuint32_t)F2Box<2>::BX, (cuuint32_t)F2Box<2>::BY}, box3[3] = {(cuuint32_t)F2Box<3>::BX, (cuuint32_t)F2Box<3>::BY, (cuuint32_t)F2Box<3>::BZ};
    const cuuint32_t es[3] = {1, 1, 1};
    if (kind == 1) { box2[0] = 32; box2[1] = 32; box3[0] = 32; box3[1] = 8; box3[2] = 4; }
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char *e = getenv("PB200_TMA_PROMO")) { const int v = atoi(e); promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B)); }
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)N, (void *)ptr, dims, strides, N == 2 ? box2 : box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(s->ctx, PB200_ECUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    F.tmaps[std::make_pair(ptr, kind)] = m;
    *out = m;
    return PB200_OK;
}
static int fold_maps(pb200_solver *s, const FVec &a, const FVec *b, F2Maps *out)
{
    memset(out, 0, sizeof(*out));
    int rc;
    for (int f = 0; f < s->F.d.nbulk; ++f) {
        if ((rc = fold_tmap(s, a.f[f], &out->a[f]))) return rc;
        if (b && (rc = fold_tmap(s, b->f[f], &out->b[f]))) return rc;
    }
    if (s->F.d.nbulk == 1) { out->a[1] = out->a[0]; out->b[1] = out->b[0]; }
    return PB200_OK;
}
template <int N, int MODE>
static int fold2_launch(pb200_solver *s, const Items &L, const F2Maps &maps, const F2Args &A, cudaStream_t st)
{
    pb200_ctx *ctx = s->ctx;
    constexpr int smem = 2 * (MODE == 5 ? 2 : 1) * F2Box<N>::SLOT;
    static bool attr_set[64] = {};   // per DEVICE: a function attribute belongs to the context of the device it was set on
    if (!attr_set[ctx->device & 63]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kf2_apply<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set[ctx->device & 63] = true; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf2_apply<N, MODE>, FCH, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 2; }
    int grid = L.n < ctx->sm_count * nb ? L.n : ctx->sm_count * nb;
    if (grid < 1) grid = 1;
    kf2_apply<N, MODE><<<grid, FCH, smem, st>>>(maps, s->g, s->F.d, L, A);
    LAUNCH_CHECK(ctx);
    return PB200_OK;
}
// the staged apply of `mode` on list L (N >= 2 only)
static int fold2_apply(pb200_solver *s, const Items &L, const F2Maps &maps, const F2Args &A, int mode, cudaStream_t st)
{
    const int N = s->g.N;
#define F2L(M_) (N == 2 ? fold2_launch<2, M_>(s, L, maps, A, st) : fold2_launch<3, M_>(s, L, maps, A, st))
    switch (mode) {
    case 0: return F2L(0);
    case 1: return F2L(1);
    case 2: return F2L(2);
    case 3: return F2L(3);
    case 4: return F2L(4);
    default: return F2L(5);
    }
#undef F2L
}

// ---- the pipelined kernel (kf3_apply) on a list of interior constant-coefficient tiles ------------------------------------------------
static int fold_maps3(pb200_solver *s, const FVec &a, const FVec *b, const FVec *t, F3Maps *out)
{
    memset(out, 0, sizeof(*out));
    int rc;
    for (int f = 0; f < s->F.d.nbulk; ++f) {
        if ((rc = fold_tmap(s, a.f[f], &out->a[f]))) return rc;
        if (b && (rc = fold_tmap(s, b->f[f], &out->b[f]))) return rc;
        if (t && (rc = fold_tmap(s, t->f[f], &out->t[f], 1))) return rc;
    }
    if (s->F.d.nbulk == 1) { out->a[1] = out->a[0]; out->b[1] = out->b[0]; out->t[1] = out->t[0]; }
    return PB200_OK;
}
// one-wave grid of the pipelined kernel (MODE 5) on list L
template <int N>
static int fold3_grid5(pb200_solver *s, const Items &L)
{
    constexpr int S = N == 2 ? 3 : 2;
    constexpr int smem = S * (2 * F2Box<N>::SLOT + FTILE * 8) + 128;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf3_apply<N, 5, S>, FCH + 32, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    int grid = L.n < s->ctx->sm_count * nb ? L.n : s->ctx->sm_count * nb;
    return grid < 1 ? 1 : grid;
}
template <int N, int BH>
static int fold3_launch_bh(pb200_solver *s, const Items &L, const F3Maps &maps, const F2Args &A, cudaStream_t st)
{
    pb200_ctx *ctx = s->ctx;
    constexpr int S = N == 2 ? 3 : 2;
    constexpr int smem = S * (2 * F2Box<N>::SLOT + FTILE * 8) + 128;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kf3_apply<N, 5, S, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set[ctx->device & 63] = true; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf3_apply<N, 5, S, BH>, FCH + 32, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    int grid = L.n < ctx->sm_count * nb ? L.n : ctx->sm_count * nb;
    if (grid < 1) grid = 1;
    kf3_apply<N, 5, S, BH><<<grid, FCH + 32, smem, st>>>(maps, L, A, 1, 0);
    LAUNCH_CHECK(ctx);
    return PB200_OK;
}
template <int N, int MODE>
static int fold3_launch(pb200_solver *s, const Items &L, const F3Maps &maps, const F2Args &A, int has_t, cudaStream_t st)
{
    pb200_ctx *ctx = s->ctx;
    constexpr int S = MODE == 5 ? (N == 2 ? 3 : 2) : (N == 2 ? 4 : 3);
    constexpr bool TT = MODE == 5 || MODE == 2 || MODE == 4;
    constexpr int STAGE = (MODE == 5 ? 2 : 1) * F2Box<N>::SLOT + (TT ? FTILE * 8 : 0);
    constexpr int smem = S * STAGE + 128;
    static bool attr_set[64] = {};   // per DEVICE (one process may drive several: pb200_init_multi)
    if (!attr_set[ctx->device & 63]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kf3_apply<N, MODE, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set[ctx->device & 63] = true; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf3_apply<N, MODE, S>, FCH + 32, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
    int grid = L.n < ctx->sm_count * nb ? L.n : ctx->sm_count * nb;
    const int dbg = getenv("PB200_DBG_F3") ? atoi(getenv("PB200_DBG_F3")) : 0;
    if ((dbg & 4) && grid > ctx->sm_count) grid = ctx->sm_count;
    if (grid < 1) grid = 1;
    if (N == 3 && MODE == 5) {
        // 3-D fused apply: ONE block per SM with a four-stage pipeline (171 KB of shared memory) instead of two blocks with two stages each -- the same
        // bytes in flight, but the boxes of neighbouring tiles then share their halo planes in L2 (ncu, 1024 x 1024 x 128: 6.8 -> 5.5 GB of DRAM reads
        // per launch, 1.53 -> 1.39 ms).  PB200_F3_S = 2 (two blocks per SM) .. 5.
        const int sreq = getenv("PB200_F3_S") ? atoi(getenv("PB200_F3_S")) : 4;
        auto go = [&](auto kern, int smemS) -> int {
            static bool attrS[64][8] = {};
            const int si = smemS / STAGE;
            if (!attrS[ctx->device & 63][si & 7]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smemS)); attrS[ctx->device & 63][si & 7] = true; }
            int nbS = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbS, kern, FCH + 32, smemS) != cudaSuccess || nbS < 1) { cudaGetLastError(); nbS = 1; }
            int gS = L.n < ctx->sm_count * nbS ? L.n : ctx->sm_count * nbS;
            if ((dbg & 4) && gS > ctx->sm_count) gS = ctx->sm_count;
            if (gS < 1) gS = 1;
            kern<<<gS, FCH + 32, smemS, st>>>(maps, L, A, has_t, dbg);
            LAUNCH_CHECK(ctx);
            return PB200_OK;
        };
        if (!(dbg & 8)) {
            if (sreq == 3) return go(kf3_apply<3, 5, 3>, 3 * STAGE + 128);
            if (sreq == 4) return go(kf3_apply<3, 5, 4>, 4 * STAGE + 128);
            if (sreq == 5) return go(kf3_apply<3, 5, 5>, 5 * STAGE + 128);
        }
    }
    if (N == 3 && MODE != 5 && getenv("PB200_F3_PLAIN_S6")) {   // experiment: plain 3-D applies with one block per SM, six stages -- measured WORSE (512^3 CN: 36.4 -> 42.7 ms per step)
        static bool attr6[64] = {};
        constexpr int smem6 = 6 * STAGE + 128;
        if (!attr6[ctx->device & 63]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kf3_apply<3, MODE, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem6)); attr6[ctx->device & 63] = true; }
        int nb6 = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb6, kf3_apply<3, MODE, 6>, FCH + 32, smem6) != cudaSuccess || nb6 < 1) { cudaGetLastError(); nb6 = 1; }
        int g6 = L.n < ctx->sm_count * nb6 ? L.n : ctx->sm_count * nb6;
        if (g6 < 1) g6 = 1;
        kf3_apply<3, MODE, 6><<<g6, FCH + 32, smem6, st>>>(maps, L, A, has_t, dbg);
        LAUNCH_CHECK(ctx);
        return PB200_OK;
    }
    if ((dbg & 8) && MODE == 5) {   // one stage: no overlap, tests the pipeline logic
        constexpr int smem1 = STAGE + 128;
        static bool attr1[64] = {};
        if (!attr1[ctx->device & 63]) { CUDA_TRY(ctx, cudaFuncSetAttribute(kf3_apply<N, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1)); attr1[ctx->device & 63] = true; }
        kf3_apply<N, MODE, 1><<<grid, FCH + 32, smem1, st>>>(maps, L, A, has_t, dbg);
    } else
    kf3_apply<N, MODE, S><<<grid, FCH + 32, smem, st>>>(maps, L, A, has_t, dbg);
    LAUNCH_CHECK(ctx);
    return PB200_OK;
}
static int fold3_apply(pb200_solver *s, const Items &L, const F3Maps &maps, const F2Args &A, int mode, int has_t, cudaStream_t st)
{
    const int N = s->g.N;
#define F3L(M_) (N == 2 ? fold3_launch<2, M_>(s, L, maps, A, has_t, st) : fold3_launch<3, M_>(s, L, maps, A, has_t, st))
    switch (mode) {
    case 0: return F3L(0);
    case 1: return F3L(1);
    case 2: return F3L(2);
    case 3: return F3L(3);
    case 4: return F3L(4);
    default: return F3L(5);
    }
#undef F3L
}

// y = M^ x with the dot products of `mode` published (dense part -> *_D slots, band part -> *_B slots) and summed over the ranks
// mode 4 (one step of the polynomial preconditioner, y = pc.r aux + pc.z x + pc.A M^ x): the partial sums of (aux, y) go to the
// caller's slots (dense part, band part) and are reduced over the ranks by the caller together with the rest of their group
static int fold_apply(pb200_solver *s, const FVec &x, const FVec &y, const FVec &aux, int mode, StopCrit stop = StopCrit{0.0, 0.0, -1},
                      PolyCoef pc = PolyCoef{0.0, 0.0, 0.0}, int slotD4 = FS_TMP, int slotB4 = FS_TMP + 1)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    FoldSys &F = s->F;
    int rc;
    if ((rc = fold_halo(s, x))) return rc;
    double *res = ctx->d_results;
    const int grid = fold_grid(s);
    double *slotD = res + (mode == 4 ? slotD4 : (mode == 3 ? FS_TS_D : FS_SIG_D)), *slotB = res + (mode == 4 ? slotB4 : (mode == 3 ? FS_TS_B : FS_SIG_B));
    prof_mark(ctx, PB_PROF_APPLY);
    ctx->apply_launches++;
    if (F.tma_ok) {   // staged tile + halo (fold2.cuh)
        F2Maps maps;
        if ((rc = fold_maps(s, x, nullptr, &maps))) return rc;
        F2Args A;
        memset(&A, 0, sizeof(A));
        A.a = x; A.y = y; A.aux = aux; A.stop = stop; A.pc = pc; A.partials = ctx->d_partials; A.results = slotD; A.counter = ctx->d_counter; A.res = res;
        A.dbg = getenv("PB200_DBG_F3") ? atoi(getenv("PB200_DBG_F3")) : 0;
        if (F.pipe && !getenv("PB200_DBG_NOPIPE_PLAIN")) {   // interior constant-coefficient tiles: the pipelined kernel; the rest adds its share of the dot products afterwards
            const int has_t = (mode == 2 || mode == 4) && aux.f[0] != x.f[0];
            F3Maps m3;
            if ((rc = fold_maps3(s, x, nullptr, has_t ? &aux : nullptr, &m3))) return rc;
            for (int p = 0; p < 2; ++p) for (int dd = 0; dd < PB_MAXD; ++dd) A.off[p][dd] = F.d.off[p][dd];
            if ((rc = fold3_apply(s, F.IA, m3, A, mode, has_t, ctx->stream))) return rc;
        } else if ((rc = fold2_apply(s, F.IA, maps, A, mode, ctx->stream))) return rc;
    } else {
#define FOLD_DENSE(M_) DISPATCH_N(g.N, (kf_apply_dense<N, M_><<<wave_grid(s, kf_apply_dense<N, M_>), FCH, 0, ctx->stream>>>(g, F.d, F.IA, x, y, aux, ctx->d_partials, slotD, ctx->d_counter, res, stop, pc)))
    if (mode == 0) FOLD_DENSE(0); else if (mode == 1) FOLD_DENSE(1); else if (mode == 2) FOLD_DENSE(2); else if (mode == 3) FOLD_DENSE(3); else FOLD_DENSE(4);
#undef FOLD_DENSE
    LAUNCH_CHECK(ctx);
    }
    prof_mark(ctx, PB_PROF_APPLY);
    if (F.d.has_w) {   // also on a rank without band cells: the kernel must refresh its partial-sum slots (to 0) before the in-place reduction
        prof_mark(ctx, PB_PROF_BAPPLY);
        struct Mark { pb200_ctx *c; ~Mark() { prof_mark(c, PB_PROF_BAPPLY); } } mark_{ctx};
        const int gb = band_wgrid(F.d.nE);
#define FOLD_BAND(M_) DISPATCH_N(g.N, (BAND_APPLY_LAUNCH(M_, g, F.d, x, y, aux, ctx->d_partials, slotB, ctx->d_counter, res, stop, pc)))
        if (mode == 0) FOLD_BAND(0); else if (mode == 1) FOLD_BAND(1); else if (mode == 2) FOLD_BAND(2); else if (mode == 3) FOLD_BAND(3); else FOLD_BAND(4);
#undef FOLD_BAND
        LAUNCH_CHECK(ctx);
    }
    if (mode != 0 && mode != 4 && (rc = allreduce_results(ctx, mode == 3 ? FS_TS_D : FS_SIG_D, mode == 3 ? 4 : 2))) return rc;
    return PB200_OK;
}

// Polynomial preconditioner of the CG on the folded system: z = q_m(M^) r, q_m = the Chebyshev polynomial of degree m for the interval
// [lo, hi] of the BULK spectrum (the few low interface modes below it are the band preconditioner's business: z = q_m(M^) r + (q_B(M^_BB) - 1)
// r_B, an SPD sum).  Why: a CG iteration moves 80 B per unknown around one 16 B operator apply; a polynomial step is the apply with an axpby
// epilogue (kf_apply_* MODE 4: 24-32 B), so the same error reduction costs ~30 % fewer bytes and (m + 1) times fewer reductions
// (tests/experiments/krylov_experiment5.py: 35 -> 14 outer iterations at m = 2).  Three-term recurrence with z_1 = r / theta never stored:
//   step 1:  z_2 = ((1 + rho_1 rho_0) / theta + 2 rho_1 / delta) r - (2 rho_1 / (delta theta)) M^ r
//   step 2:  z_3 = (1 + rho_2 rho_1) z_2 + (2 rho_2 / delta - rho_2 rho_1 / theta) r - (2 rho_2 / delta) M^ z_2
static void poly_coefs(double lo, double hi, PolyCoef c[2])
{
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    const double rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0), rho2 = 1.0 / (2.0 * sigma - rho1);
    c[0].r = (1.0 + rho1 * rho0) / theta + 2.0 * rho1 / delta; c[0].z = 0.0; c[0].A = -2.0 * rho1 / (delta * theta);
    c[1].r = 2.0 * rho2 / delta - rho2 * rho1 / theta; c[1].z = 1.0 + rho2 * rho1; c[1].A = -2.0 * rho2 / delta;
}
// host-only: the coefficients of the two polynomial steps, out = {r1, z1, A1, r2, z2, A2} (unit-tested against a NumPy Chebyshev iteration)
extern "C" int pb200_poly_coefs(double lo, double hi, double out[6])
{
    if (!out || !(lo > 0.0) || !(hi > lo)) return set_err(nullptr, PB200_EINVAL, "need 0 < lo < hi");
    PolyCoef c[2];
    poly_coefs(lo, hi, c);
    out[0] = c[0].r; out[1] = c[0].z; out[2] = c[0].A; out[3] = c[1].r; out[4] = c[1].z; out[5] = c[1].A;
    return PB200_OK;
}
// z = q_m(M^) r into `out` (m = 1: one step; m = 2: through F.z), (r, z) published into res[slot] (dense) and res[slot + 3] (band part)
static int fold_poly(pb200_solver *s, const FVec &r, const FVec &out, int slot, StopCrit stop)
{
    FoldSys &F = s->F;
    PolyCoef c[2];
    poly_coefs(F.poly_lo, F.poly_hi, c);
    int rc;
    if (F.poly_m == 1) return fold_apply(s, r, out, r, 4, stop, c[0], slot, slot + 3);
    if ((rc = fold_apply(s, r, F.z, r, 4, stop, c[0], FS_TMP, FS_TMP + 1))) return rc;
    return fold_apply(s, F.z, out, r, 4, stop, c[1], slot, slot + 3);
}

// set-up of the polynomial preconditioner, after fold_build (collective: every rank calls it)
static int fold_poly_setup(pb200_solver *s)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    FoldSys &F = s->F;
    int rc;
    // polynomial preconditioner (fold_poly): Chebyshev interval = bulk spectrum [1 - R, max(1 + R, lambda_max(M^))], R = 2 sum_d |c_d| of the
    // constant-coefficient interior stencil (its eigenvalues are 1 + 2 sum_d c_d cos k_d); lambda_max by a power iteration on M^ (collective)
    F.poly_m = 0;
    {
        const char *e = getenv("PB200_POLY");
        int m = e ? atoi(e) : PB200_POLY_DEFAULT;
        if (m > 2) m = 2;
        // Only without interface unknowns (monophasic problems with a Dirichlet interface).  Measured on the diphasic configs[1] system
        // (2048^2, extrapolated initial guess; iterations per step): band preconditioner alone 8.35 (what runs), no preconditioner 27.6,
        // polynomial alone 12.8, the sum q(M^) + (q_B(M^_BB) - 1) ~230 -- it is not positive definite: for band modes at the upper end of
        // the spectrum q_B - 1 ~ -0.9 outweighs q ~ 0.3 --, the symmetric product B q(M^) B 32 -- q_B^2 crushes those same modes
        // (lambda q_B^2 q ~ 0.01 against ~2-4 elsewhere).  The residual left by the extrapolated guess lives on the interface band, which
        // is why the O(band) preconditioner does more there than the bulk polynomial; combining them needs a deflation-type coupling.
        if (F.d.has_w) m = 0;
        if (m > 0) {
            double R = 0.0, cnt = 0.0;
            if (F.nitems > 0) {
                std::vector<unsigned char> hu(F.nitems);
                std::vector<double> hc((size_t)F.nitems * PB_MAXD);
                std::vector<TileRec> hr(F.nitems);
                CUDA_TRY(ctx, cudaMemcpy(hu.data(), F.uni, (size_t)F.nitems, cudaMemcpyDeviceToHost));
                CUDA_TRY(ctx, cudaMemcpy(hc.data(), F.ucoef, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
                CUDA_TRY(ctx, cudaMemcpy(hr.data(), F.rec, sizeof(TileRec) * (size_t)F.nitems, cudaMemcpyDeviceToHost));
                for (int i = 0; i < F.nitems; ++i)
                    if (hr[i].f < 2 && (hu[i] & 1) && hr[i].full) {
                        double r = 0.0;
                        for (int dd = 0; dd < g.N; ++dd) r += 2.0 * fabs(hc[(size_t)i * PB_MAXD + dd]);
                        if (r > R) R = r;
                        cnt = 1.0;
                    }
            }
            // the same numbers on every rank: mean of the ranks that hold interior tiles (they agree on uniform grids)
            double h2[2] = {cnt > 0.0 ? R : 0.0, cnt};
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_results + SL_TMP, h2, sizeof(h2), cudaMemcpyHostToDevice, ctx->stream));
            if ((rc = allreduce_results(ctx, SL_TMP, 2))) return rc;
            if ((rc = fetch_results(ctx, SL_TMP, 2, h2))) return rc;
            if (h2[1] > 0.5 && h2[0] / h2[1] < 0.95) {
                R = h2[0] / h2[1];
                // power iteration, no normalisation (lambda_max < ~2: 1.8^32 is harmless in fp64): lambda ~ (y, y) / (y, x), y = M^ x
                const int gz = wave_grid(s, kf_seed);
                kf_seed<<<gz, FCH, 0, ctx->stream>>>(F.I, F.p); LAUNCH_CHECK(ctx);
                double lam = 1.0 + R;
                FVec *a = &F.p, *b = &F.v;
                for (int it = 0; it < 32; ++it) {
                    if ((rc = fold_apply(s, *a, *b, *b, 3))) return rc;
                    std::swap(a, b);
                }
                double t4[4];
                if ((rc = fetch_results(ctx, FS_TS_D, 4, t4))) return rc;   // FS_TS_D, FS_TT_D, FS_TS_B, FS_TT_B of the last apply
                const double yx = t4[0] + t4[2], yy = t4[1] + t4[3];
                if (yx > 0.0 && yy > 0.0 && yy / yx > lam) lam = yy / yx;
                kf_zero<<<gz, FCH, 0, ctx->stream>>>(F.I, F.p); LAUNCH_CHECK(ctx);
                kf_zero<<<gz, FCH, 0, ctx->stream>>>(F.I, F.v); LAUNCH_CHECK(ctx);
                F.poly_lo = 1.0 - R;
                F.poly_hi = 1.03 * lam;
                F.poly_m = m;
                if (!F.have_z) { if ((rc = fold_alloc_vec(s, &F.z))) return rc; F.have_z = true; }
                if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] polynomial preconditioner: degree %d on [%.4f, %.4f] (R = %.4f, lambda_max ~ %.4f)\n", m, F.poly_lo, F.poly_hi, R, lam);
            }
        }
    }
    return PB200_OK;
}

#include "mg.cuh"

// Krylov solve of the folded system.  In: s->b (reference rows, known parts eliminated), s->x (initial guess on the free sets).
// Out: s->x.  The stopping test ||r^|| <= max(rtol ||b^||, atol) is on the block-Jacobi-scaled residual.
static int fold_solve(pb200_solver *s, int method, const pb200_krylov_opts &o, const GuessSpec gsp[3], bool dense_done, int *iters, int *conv, double *rnorm_out,
                      double *bnorm_out)
{
    pb200_ctx *ctx = s->ctx;
    FoldSys &F = s->F;
    const Items &I = F.I;
    int rc;
    if (method == PB200_KRYLOV_BICGSTAB && !F.have_bicg) {
        FVec *vs[] = {&F.r0, &F.s, &F.t};
        for (FVec *v : vs) if ((rc = fold_alloc_vec(s, v))) return rc;
        F.have_bicg = true;
    }
    double *res = ctx->d_results;
    CUDA_TRY(ctx, cudaMemsetAsync(res, 0, sizeof(double) * RED_SLOTS, ctx->stream));
    const int grid = fold_grid(s);
    const int gb = band_grid(F.d.nBown);
    const bool band = F.d.has_w && F.d.nBown > 0;
    // b^ and x^0
    // (dense_done: kf_rhs_dense has written b^ and x^0 on the active tiles; only the rows whose known part came afterwards are refreshed)
    if (!dense_done) { kf_to_scaled_dense<<<wave_grid(s, kf_to_scaled_dense), FCH, 0, ctx->stream>>>(F.d, I, s->b, F.b); LAUNCH_CHECK(ctx); }
    else if (s->nK > 0) { kf_to_scaled_list<<<(s->nK + 127) / 128, 128, 0, ctx->stream>>>(F.d, s->Kcell, s->nK, s->b, F.b); LAUNCH_CHECK(ctx); }
    if (band) { kf_to_scaled_band<<<gb, 128, 0, ctx->stream>>>(F.d, s->b, F.b); LAUNCH_CHECK(ctx); }
    const bool cg = method == PB200_KRYLOV_CG;
    static_assert(FS_RR0 == FS_BB + 1, "kf_resid publishes the pair (bb, rr0)");
    if (o.precond == PB200_PRECOND_MG) {   // multigrid-preconditioned CG (mg.cuh), zero initial guess
        if (!cg) return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner is a CG preconditioner");
        const ApplyCoef acm = {F.key[0], F.key[1], F.key[2], 0, 1};
        if ((!s->mg || !s->mg->ready) && (rc = mg_setup(s, acm))) { mg_free(s); return rc; }
        int itm = 0, cvm = 0;
        double rn = 0.0, bn = 0.0;
        if ((rc = mg_pcg(s, o, &itm, &cvm, &rn, &bn))) return rc;
        prof_mark(ctx, PB_PROF_EPILOGUE);
        kf_from_scaled_dense<<<wave_grid(s, kf_from_scaled_dense), FCH, 0, ctx->stream>>>(F.d, I, F.x, s->x); LAUNCH_CHECK(ctx);
        prof_mark(ctx, PB_PROF_EPILOGUE);
        *iters = itm; *conv = cvm; *rnorm_out = rn; *bnorm_out = bn;
        return PB200_OK;
    }
    // fused CG iteration (fold2.cuh): p update + apply in one TMA-staged kernel, ghost-class tiles and the halo exchange on a second stream
    const bool fused = cg && F.tma_ok && !getenv("PB200_NO_FUSED");
    if (fused && !F.have_p2) { if ((rc = fold_alloc_vec(s, &F.p2))) return rc; F.have_p2 = true; }
    if (fused && F.poly_m > 0 && !F.have_zz) { if ((rc = fold_alloc_vec(s, &F.zz))) return rc; F.have_zz = true; }
    const int pzero = fused ? 1 : 0;
    const bool warm = gsp[0].m > 0;
    if (warm) {
        if (!dense_done) { kf_guess_dense<<<wave_grid(s, kf_guess_dense), FCH, 0, ctx->stream>>>(F.d, I, gsp[0], gsp[1], F.x); LAUNCH_CHECK(ctx); }
        if (band) { kf_guess_band<<<gb, 128, 0, ctx->stream>>>(F.d, gsp[0], gsp[1], gsp[2], F.x); LAUNCH_CHECK(ctx); }
        if ((rc = fold_apply(s, F.x, F.v, F.v, 0))) return rc;
        kf_resid<<<wave_grid(s, kf_resid), FCH, 0, ctx->stream>>>(I, F.b, F.v, 1, F.r, F.p, F.r0, cg ? 0 : 1, pzero, ctx->d_partials, res + FS_BB, ctx->d_counter); LAUNCH_CHECK(ctx);
    } else {
        kf_zero<<<grid, FCH, 0, ctx->stream>>>(I, F.x); LAUNCH_CHECK(ctx);
        kf_resid<<<wave_grid(s, kf_resid), FCH, 0, ctx->stream>>>(I, F.b, F.v, 0, F.r, F.p, F.r0, cg ? 0 : 1, pzero, ctx->d_partials, res + FS_BB, ctx->d_counter); LAUNCH_CHECK(ctx);
    }
    if ((rc = allreduce_results(ctx, FS_BB, 2))) return rc;
    // No host look at ||b||, ||r0|| here: the device-side stopping test (fold_done) needs neither, and a converged start simply turns
    // every kernel of the first chunk into a no-op.  bnorm / rnorm / iteration count come back with the first fetch.
    double bnorm = 0.0, rnorm = 0.0, tol = 0.0;
    int it = 0, converged = 0;
    int cur = 0;
    {
        // (rho, rr) pair 0 = (rr0, rr0)
        CUDA_TRY(ctx, cudaMemcpyAsync(res + FS_PAIR0, res + FS_RR0, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(res + FS_PAIR0 + 1, res + FS_RR0, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        const bool prec = cg && F.d.has_w && F.prec;   // F.prec is the same on every rank
        const int gE = band_wgrid(F.d.nE);
        const StopCrit nostop = {0.0, 0.0, -1};
        const bool poly = cg && F.poly_m > 0;
        if (poly && (rc = fold_poly(s, F.r, fused ? F.zz : F.p, FS_PAIR0, nostop))) return rc;   // p0 = q(M^) r0 (kf_resid had set p0 = r0; fused: z0, p0 is formed by the first apply)
        if (cg) {
            if (prec) {   // p0 = z0 = r0 + (q(M^_BB) - 1) r0_B ; rho0 = (r0, z0)
                DISPATCH_N(s->g.N, (BAND_POLY_LAUNCH(s->g, F.d, F.r, F.dz, F.pa0 - 1.0, F.pa1, ctx->d_partials, res + FS_PAIR0 + 2, ctx->d_counter, res, nostop)));
                LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, FS_PAIR0 + 2, 1))) return rc;
                if (!fused) { kf_band_put<<<gb, 128, 0, ctx->stream>>>(F.d, F.p, F.dz, 1.0, 1, res, nostop); LAUNCH_CHECK(ctx); }
            }
            if (poly && ((rc = allreduce_results(ctx, FS_PAIR0, 1)) || (rc = allreduce_results(ctx, FS_PAIR0 + 3, 1)))) return rc;   // (slots 1, 2 are global already)
        }
        // The kernels test the residual themselves (fold_done) and fall through once it is below the tolerance, so `check_every`
        // iterations are queued between two host looks; FS_ITERS counts the iterations that really ran.
        // band heads (fold2.cuh): the two O(band) launches of the iteration folded into the streaming kernels -- one rank, pipelined kernel, band preconditioner,
        // SMALL bands (2-D problems: the separate launches are pure latency there; on the large bands of 3-D problems they are bandwidth work and were
        // measured equal either way -- 1024 x 1024 x 128 diphasic: 68.3 vs 68.2 ms per step)
        const bool bandfuse = cg && fused && prec && F.bandfuse_ok && F.pipe && !getenv("PB200_DBG_NOPIPE_FUSED") && !getenv("PB200_NO_BANDFUSE");   // (the same on every rank)
        if (bandfuse && !F.have_r2) {
            CUDA_TRY(ctx, cudaMalloc((void **)&F.r2.f[2], sizeof(double) * (size_t)(F.d.nB > 0 ? F.d.nB : 1)));
            CUDA_TRY(ctx, cudaMalloc((void **)&F.rE, sizeof(double) * 4 * (size_t)F.d.nEp));
            F.have_r2 = true;
        }
        if (bandfuse) { F.r2.f[0] = F.r.f[0]; F.r2.f[1] = F.r.f[1]; }
        BandHead bhd;
        memset(&bhd, 0, sizeof(bhd));
        if (bandfuse) {
            bhd.on = 1; bhd.nE = F.d.nE; bhd.nEp = F.d.nEp; bhd.nB = F.d.nB; bhd.nbulk = F.d.nbulk;
            bhd.Ecell = F.d.Ecell; bhd.EB = F.d.EB; bhd.EnbrB = F.d.EnbrB; bhd.EnbrE = F.d.EnbrE; bhd.EofB = F.d.EofB; bhd.eord = F.d.eord; bhd.Eblk = F.d.Eblk;
            bhd.ya = F.d.ya; bhd.dzw = F.dz; bhd.ld0 = F.d.ld0; bhd.dP = F.d.dP;
            for (int dd = 0; dd < PB_MAXD; ++dd) bhd.sq[dd] = F.d.sq[dd];
            bhd.ca = F.pa0 - 1.0; bhd.cb = F.pa1;
            bhd.Bq = F.d.Bq; bhd.Bidx = F.d.Bidx; bhd.Bblk = F.d.Bblk; bhd.nBp = F.d.nBp;
            if (F.d.nE > 0) { kf2_gather_rE<<<(F.d.nE + 255) / 256, 256, 0, ctx->stream>>>(bhd, F.r, F.rE, F.rE + 2 * (size_t)F.d.nEp); LAUNCH_CHECK(ctx); }
        }
        auto enqueue = [&](int curp) -> int {   // one Krylov iteration reading pair `curp`, publishing pair curp ^ 1
            const int nxt = curp ^ 1;
            StopCrit st = {o.rtol * o.rtol, o.atol * o.atol, FS_TRIPLE(curp) + 1};
            StopCrit stn = {o.rtol * o.rtol, o.atol * o.atol, FS_TRIPLE(nxt) + 1};
            int rc2;
            if (cg && fused) {
                const bool multi = ctx->nranks > 1;
                const FVec &pold = curp ? F.p2 : F.p, &pnew = curp ? F.p : F.p2;    // iteration j reads P[j & 1], writes P[(j + 1) & 1]
                const FVec &rold = bandfuse && curp ? F.r2 : F.r, &rnew = bandfuse && !curp ? F.r2 : F.r;   // band heads: r double-buffered like p
                const FVec &zsrc = poly ? F.zz : rold;
                // The second stream exists to overlap the HALO EXCHANGE with the interior tiles.  On one rank the side work is the pointwise update of
                // the interface unknowns only, and running it beside the staged kernel cost more than it hid (512^3 diphasic: 85 vs 65 ms per step).
                cudaStream_t st2 = (ctx->profile || !multi || getenv("PB200_DBG_SERIAL")) && !getenv("PB200_DBG_FORK") ? ctx->stream : ctx->stream2;
                const Items &Lpupd = bandfuse ? F.IG1nw : F.IG1;      // (band heads update the interface unknowns themselves)
                const bool side = Lpupd.n > 0 || multi;
                F2Args A;
                memset(&A, 0, sizeof(A));
                A.bh = bhd;
                A.a = zsrc; A.pold = pold; A.y = F.v; A.pnew = pnew; A.xs = F.x; A.aux = F.v;
                A.dz = prec ? F.dz : nullptr; A.bord = F.bord; A.nB = F.d.nB;
                A.sl_old = FS_TRIPLE(nxt); A.sl_cur = FS_TRIPLE(curp); A.stop = st; A.res = res;
                A.partials = ctx->d_partials; A.counter = ctx->d_counter; A.results = res + FS_SIG_D;
                A.dbg = getenv("PB200_DBG_F3") ? atoi(getenv("PB200_DBG_F3")) : 0;
                A.l2hint = getenv("PB200_L2HINT") ? atoi(getenv("PB200_L2HINT")) : 0;
                for (int pp = 0; pp < 2; ++pp) for (int dd = 0; dd < PB_MAXD; ++dd) A.off[pp][dd] = F.d.off[pp][dd];
                if (side) {
                    if (st2 != ctx->stream) { CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream)); CUDA_TRY(ctx, cudaStreamWaitEvent(st2, ctx->ev_fork, 0)); }
                    if (Lpupd.n > 0) {   // ghost-class tiles and the compact interface unknowns: p_k and x pointwise
                        prof_mark(ctx, PB_PROF_PUPD);
                        int g1 = Lpupd.n < ctx->sm_count * 4 ? Lpupd.n : ctx->sm_count * 4;
                        kf2_pupd<<<g1, FCH, 0, st2>>>(Lpupd, A); LAUNCH_CHECK(ctx);
                        prof_mark(ctx, PB_PROF_PUPD);
                    }
                    if (multi && (rc2 = fold_halo(s, pnew, st2))) return rc2;   // ghost planes of p_k: in flight while the interior class computes
                }
                F2Maps mi;
                if ((rc2 = fold_maps(s, zsrc, &pold, &mi))) return rc2;
                const bool split = F.pipe && !getenv("PB200_DBG_NOPIPE_FUSED");     // interior class: pipelined kernel on the constant-coefficient tiles + general kernel on the rest (beside it)
                prof_mark(ctx, PB_PROF_APPLY);
                ctx->apply_launches++;
                if (bandfuse) {   // the same launch with the apply head
                    F3Maps m3;
                    if ((rc2 = fold_maps3(s, zsrc, &pold, &F.x, &m3))) return rc2;
                    const int N3 = s->g.N;
                    const Items &L5 = multi ? F.IAi_all : F.IA;
                    const int g5 = N3 == 2 ? fold3_grid5<2>(s, L5) : fold3_grid5<3>(s, L5);
                    const bool two_lanes = (long long)F.d.nE * 2 <= (long long)g5 * FCH;
                    if (N3 == 2) rc2 = two_lanes ? fold3_launch_bh<2, 2>(s, L5, m3, A, ctx->stream) : fold3_launch_bh<2, 1>(s, L5, m3, A, ctx->stream);
                    else rc2 = two_lanes ? fold3_launch_bh<3, 2>(s, L5, m3, A, ctx->stream) : fold3_launch_bh<3, 1>(s, L5, m3, A, ctx->stream);
                    if (rc2) return rc2;
                } else if (split) {   // one pipelined launch: constant-coefficient interior tiles and general tiles alike
                    F3Maps m3;
                    if ((rc2 = fold_maps3(s, zsrc, &pold, &F.x, &m3))) return rc2;
                    if ((rc2 = fold3_apply(s, multi ? F.IAi_all : F.IA, m3, A, 5, 1, ctx->stream))) return rc2;
                } else if ((rc2 = fold2_apply(s, multi ? F.IAi_all : F.IA, mi, A, 5, ctx->stream))) return rc2;
                prof_mark(ctx, PB_PROF_APPLY);
                if (side && st2 != ctx->stream) { CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, st2)); CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0)); }
                if (multi) {
                    // ghost-class tiles: plain staged apply of p_k.  After the join: their boxes also read p_k cells of interior-class tiles,
                    // which the fused kernels have only now finished writing.
                    F2Maps mg;
                    if ((rc2 = fold_maps(s, pnew, nullptr, &mg))) return rc2;
                    F2Args G = A;
                    G.a = pnew; G.results = res + FS_SIG_G;
                    prof_mark(ctx, PB_PROF_APPLY);
                    if (split) {
                        F3Maps m3g;
                        if ((rc2 = fold_maps3(s, pnew, nullptr, nullptr, &m3g))) return rc2;
                        if ((rc2 = fold3_apply(s, F.IAg, m3g, G, 1, 0, ctx->stream))) return rc2;
                    } else if ((rc2 = fold2_apply(s, F.IAg, mg, G, 1, ctx->stream))) return rc2;
                    prof_mark(ctx, PB_PROF_APPLY);
                }
                if (getenv("PB200_DBG_CHECK") && !multi) {
                    // recompute v = M^ p_k with the register kernel and compare, list by list
                    if (!F.have_bicg) { FVec *vs[] = {&F.r0, &F.s, &F.t}; for (FVec *vv : vs) if ((rc2 = fold_alloc_vec(s, vv))) return rc2; F.have_bicg = true; }
                    const StopCrit ns = {0.0, 0.0, -1};
                    DISPATCH_N(s->g.N, (kf_apply_dense<N, 0><<<wave_grid(s, kf_apply_dense<N, 0>), FCH, 0, ctx->stream>>>(s->g, F.d, F.IA, pnew, F.t, F.t, ctx->d_partials, res + FS_TMP, ctx->d_counter, res, ns, PolyCoef{0.0, 0.0, 0.0})));
                    LAUNCH_CHECK(ctx);
                    double *d_out = nullptr;
                    CUDA_TRY(ctx, cudaMalloc((void **)&d_out, 8 * sizeof(double)));
                    const Items *Ls[2] = {&F.IAgen, &F.IAf};
                    for (int q = 0; q < 2; ++q) {
                        CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, 8 * sizeof(double), ctx->stream));
                        { const int big = 1 << 30; CUDA_TRY(ctx, cudaMemcpyAsync(d_out + 4, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); }
                        if (Ls[q]->n > 0) { kf2_dbg_compare<<<Ls[q]->n < 1024 ? Ls[q]->n : 1024, FCH, 0, ctx->stream>>>(*Ls[q], F.v, F.t, 1e-11, d_out); LAUNCH_CHECK(ctx); }
                        double h[8];
                        CUDA_TRY(ctx, cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
                        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                        double done_h[2];
                        CUDA_TRY(ctx, cudaMemcpy(done_h, res + st.sl_rr, sizeof(double), cudaMemcpyDeviceToHost));
                        fprintf(stderr, "[pb200] check v (%s list, %d tiles): %g cells differ, last item %g flags %g max diff %.3e; items %d..%d, tiles (first cell bad) %g\n", q ? "fast" : "general", Ls[q]->n, h[0], h[1], h[2], h[3], *(int *)&h[4], *(int *)&h[5], h[6]);
                    }
                    cudaFree(d_out);
                    double dbg2[3];
                    CUDA_TRY(ctx, cudaMemcpy(dbg2, res + 29, sizeof(dbg2), cudaMemcpyDeviceToHost));
                    { double bad; CUDA_TRY(ctx, cudaMemcpy(&bad, res + 28, sizeof(double), cudaMemcpyDeviceToHost)); if (bad != 0.0) fprintf(stderr, "[pb200] BOUNDS violation code %g\n", bad); }
                    fprintf(stderr, "[pb200] check kf3: header mismatches %g, staged-box mismatches %g, v vs global recompute mismatches %g (cumulative)\n", dbg2[1], dbg2[2], dbg2[0]);
                }
                if (bandfuse) {
                    if (multi && (rc2 = allreduce_results(ctx, FS_SIG_D, 3))) return rc2;
                    prof_mark(ctx, PB_PROF_UPDATE);
                    // head blocks (small bands only: see bandfuse): an eighth of the wave, 8 lanes per band cell
                    const int gw = s->g.N == 2 ? wave_grid(s, kf2_update_b<2, 8>) : wave_grid(s, kf2_update_b<3, 8>);
                    int HB = getenv("PB200_BANDFUSE_HB") ? atoi(getenv("PB200_BANDFUSE_HB")) : gw / 8;
                    if (HB > gw / 2) HB = gw / 2;
                    if (gw < 8) HB = 0;
#define K2B(N_) kf2_update_b<N_, 8><<<gw, FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), F.v, rold, rnew, F.rE + (curp ? 2 : 0) * (size_t)F.d.nEp, F.rE + (curp ? 0 : 2) * (size_t)F.d.nEp, pnew, bhd, HB, multi ? 0 : 1, ctx->d_partials, ctx->d_counter, st)
                    if (s->g.N == 2) K2B(2); else K2B(3);
#undef K2B
                    LAUNCH_CHECK(ctx);
                    prof_mark(ctx, PB_PROF_UPDATE);
                    if (multi) {
                        if ((rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), 3))) return rc2;
                        kf2_carry<<<1, 32, 0, ctx->stream>>>(res, FS_TRIPLE(curp), FS_TRIPLE(nxt), st); LAUNCH_CHECK(ctx);
                    }
                    return PB200_OK;
                }
                if (F.d.has_w) {     // band part of v = M^ p_k (needs p_k everywhere)
                    prof_mark(ctx, PB_PROF_BAPPLY);
                    DISPATCH_N(s->g.N, (BAND_APPLY_LAUNCH(1, s->g, F.d, pnew, F.v, F.v, ctx->d_partials, res + FS_SIG_B, ctx->d_counter, res, st, PolyCoef{0.0, 0.0, 0.0},
                                                                                                          (split && prec) ? F.dz : nullptr)));
                    LAUNCH_CHECK(ctx);
                    prof_mark(ctx, PB_PROF_BAPPLY);
                }
                if (multi && (rc2 = allreduce_results(ctx, FS_SIG_D, 3))) return rc2;
                prof_mark(ctx, PB_PROF_UPDATE);
                kf2_update<<<wave_grid(s, kf2_update), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), multi ? 0 : 1, F.v, F.r, ctx->d_partials, ctx->d_counter, st); LAUNCH_CHECK(ctx);
                prof_mark(ctx, PB_PROF_UPDATE);
                if (poly && (rc2 = fold_poly(s, F.r, F.zz, FS_TRIPLE(nxt), st))) return rc2;
                if (!prec && (rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), poly ? FS_NGROUP : 2))) return rc2;
                if (prec) {
                    prof_mark(ctx, PB_PROF_BPREC);
                    DISPATCH_N(s->g.N, (BAND_POLY_LAUNCH(s->g, F.d, F.r, F.dz, F.pa0 - 1.0, F.pa1, ctx->d_partials, res + FS_TRIPLE(nxt) + 2, ctx->d_counter, res, st,
                                                                                   split ? pnew : FVec{{nullptr, nullptr, nullptr}})));
                    LAUNCH_CHECK(ctx);
                    prof_mark(ctx, PB_PROF_BPREC);
                    if ((rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), poly ? FS_NGROUP : 3))) return rc2;
                }
                if (multi) { kf2_carry<<<1, 32, 0, ctx->stream>>>(res, FS_TRIPLE(curp), FS_TRIPLE(nxt), st); LAUNCH_CHECK(ctx); }
                return PB200_OK;
            }
            if (cg) {
                if ((rc2 = fold_apply(s, F.p, F.v, F.v, 1, st))) return rc2;
                prof_mark(ctx, PB_PROF_UPDATE);
                kf_cg_update<<<wave_grid(s, kf_cg_update), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), F.v, F.r, ctx->d_partials, ctx->d_counter, st); LAUNCH_CHECK(ctx);
                prof_mark(ctx, PB_PROF_UPDATE);
                // z = q(M^) r (polynomial preconditioner, into F.v: free until the next apply) -- rho_new = (r, q(M^) r) replaces (r, r)
                if (poly && (rc2 = fold_poly(s, F.r, F.v, FS_TRIPLE(nxt), st))) return rc2;
                if (!prec && (rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), poly ? FS_NGROUP : 2))) return rc2;
                if (prec) {   // z += (q(M^_BB) - 1) r_B on the band: rho_new += (r_B, dz_B)
                    prof_mark(ctx, PB_PROF_BPREC);
                    DISPATCH_N(s->g.N, (BAND_POLY_LAUNCH(s->g, F.d, F.r, F.dz, F.pa0 - 1.0, F.pa1, ctx->d_partials, res + FS_TRIPLE(nxt) + 2, ctx->d_counter, res, st)));
                    LAUNCH_CHECK(ctx);
                    prof_mark(ctx, PB_PROF_BPREC);
                    if ((rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), poly ? FS_NGROUP : 3))) return rc2;
                }
                prof_mark(ctx, PB_PROF_PUPD);
                kf_cg_p<<<wave_grid(s, kf_cg_p), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), poly ? F.v : F.r, F.p, F.x, prec ? F.dz : nullptr, F.bord, F.d.nB, st, stn); LAUNCH_CHECK(ctx);
                prof_mark(ctx, PB_PROF_PUPD);
            } else {
                if ((rc2 = fold_apply(s, F.p, F.v, F.r0, 2, st))) return rc2;
                kf_bicg_s<<<wave_grid(s, kf_bicg_s), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), F.r, F.v, F.s, st); LAUNCH_CHECK(ctx);
                if ((rc2 = fold_apply(s, F.s, F.t, F.t, 3, st))) return rc2;
                kf_bicg_xr<<<wave_grid(s, kf_bicg_xr), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), F.p, F.s, F.t, F.r0, F.x, F.r, ctx->d_partials, ctx->d_counter, st); LAUNCH_CHECK(ctx);
                if ((rc2 = allreduce_results(ctx, FS_TRIPLE(nxt), 2))) return rc2;
                kf_bicg_p<<<wave_grid(s, kf_bicg_p), FCH, 0, ctx->stream>>>(I, res, FS_TRIPLE(curp), FS_TRIPLE(nxt), F.r, F.v, F.p, stn); LAUNCH_CHECK(ctx);
            }
            // a skipped iteration publishes nothing: carry the converged pair over so that the next iteration sees it too (CG: inside kf_cg_p)
            if (!cg) { kf_carry_pair<<<1, 32, 0, ctx->stream>>>(res, FS_TRIPLE(curp), FS_TRIPLE(nxt), st); LAUNCH_CHECK(ctx); }
            return PB200_OK;
        };
        // Chunks of iterations between two host looks at the residual: the first chunk is sized by the iteration count of the previous
        // solve (time steps resemble each other), later ones are short.  Single GPU, no per-launch profiling: a chunk is replayed as ONE
        // CUDA graph (captured once per chunk length and parameter set), which removes the per-launch CPU cost and most inter-kernel gaps.
        // (with several ranks graphs need the peer-memory exchange: NCCL nodes inside a captured graph ran 8x slower)
        const bool use_graph = (ctx->nranks == 1 || (ctx->p2p && ctx->p2p->on) || getenv("PB200_GRAPH_NCCL")) && !ctx->profile && !getenv("PB200_NO_GRAPH") && o.maxit >= 2;
        auto even_up = [](int v) { return v < 2 ? 2 : (v + 1) & ~1; };
        const int first_chunk = even_up(F.last_iters > 0 ? F.last_iters + 1 : o.check_every);
        const int later_chunk = even_up(o.check_every < 4 ? o.check_every : 4);
        const double gkey[5] = {(double)method + 16.0 * ctx->p2p_gen + 1024.0 * F.poly_m + (fused ? 4096.0 : 0.0) + (bandfuse ? 8192.0 : 0.0), o.rtol + F.poly_lo, o.atol + F.poly_hi, prec ? F.pa0 : 0.0, prec ? F.pa1 : 0.0};
        if (use_graph && memcmp(gkey, F.graph_key, sizeof(gkey)) != 0) {
            for (auto &kv : F.graphs) cudaGraphExecDestroy(kv.second.exec);
            F.graphs.clear();
            memcpy(F.graph_key, gkey, sizeof(gkey));
        }
        int queued = 0;
        while (queued < o.maxit) {
            int chunk = queued == 0 ? first_chunk : later_chunk;
            if (chunk > o.maxit - queued) chunk = even_up(o.maxit - queued);
            if (use_graph) {
                auto gi = F.graphs.find(chunk);
                if (gi == F.graphs.end()) {
                    cudaGraph_t gr = nullptr;
                    if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] capturing a graph of %d iterations (%zu cached)\n", chunk, F.graphs.size());
                    const int64_t l0 = ctx->launches, a0 = ctx->apply_launches;
                    CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                    int rcc = PB200_OK;
                    for (int q = 0; q < chunk && !rcc; ++q) rcc = enqueue(q & 1);   // even length: starts and ends on pair 0
                    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &gr);
                    if (rcc) { if (gr) cudaGraphDestroy(gr); return rcc; }
                    CUDA_TRY(ctx, ce);
                    FoldGraph fg;
                    CUDA_TRY(ctx, cudaGraphInstantiate(&fg.exec, gr, 0));
                    cudaGraphDestroy(gr);
                    fg.launches = ctx->launches - l0; fg.applies = ctx->apply_launches - a0;
                    ctx->launches = l0; ctx->apply_launches = a0;   // capturing launches nothing
                    gi = F.graphs.emplace(chunk, fg).first;
                }
                CUDA_TRY(ctx, cudaGraphLaunch(gi->second.exec, ctx->stream));
                ctx->launches += gi->second.launches; ctx->apply_launches += gi->second.applies;
            } else {
                for (int q = 0; q < chunk; ++q) { if ((rc = enqueue(cur))) return rc; cur ^= 1; }
            }
            queued += chunk;
            double all[FS_TMP + 1];
            if ((rc = fetch_results(ctx, 0, FS_TMP + 1, all))) return rc;   // one look: rr, iteration count, ||b||^2
            if ((rc = p2p_check(ctx))) return rc;
            bnorm = sqrt(all[FS_BB]);
            tol = fmax(o.rtol * bnorm, o.atol);
            rnorm = sqrt(all[FS_TRIPLE(cur) + 1]);
            it = (int)(all[FS_ITERS] + 0.5);
            if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] queued %d it %d rnorm %.3e tol %.3e\n", queued, it, rnorm, tol);
            if (rnorm <= tol) { converged = 1; break; }
            if (!(rnorm == rnorm)) break;
        }
        F.last_iters = it;
        if (fused) {   // x += alpha_k p_k of the last iteration (the next fused apply would have added it)
            kf2_xflush<<<wave_grid(s, kf2_xflush), FCH, 0, ctx->stream>>>(I, res, F.p, F.p2, F.x); LAUNCH_CHECK(ctx);
        }
    }
    prof_mark(ctx, PB_PROF_EPILOGUE);
    kf_from_scaled_dense<<<wave_grid(s, kf_from_scaled_dense), FCH, 0, ctx->stream>>>(F.d, I, F.x, s->x); LAUNCH_CHECK(ctx);
    if (band) { kf_from_scaled_band<<<gb, 128, 0, ctx->stream>>>(F.d, F.x, s->x); LAUNCH_CHECK(ctx); }
    prof_mark(ctx, PB_PROF_EPILOGUE);
    *iters = it; *conv = converged; *rnorm_out = rnorm; *bnorm_out = bnorm;
    return PB200_OK;
}

static int build_masks(pb200_solver *s)
{
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    const int grid = sgrid(ctx, g.nown);
    const double *ct1 = s->o1->cap->ct, *ct2 = s->o2 ? s->o2->cap->ct : s->o1->cap->ct;
    // a HIGH-side Periodic row equals its LOW-side partner (same dimension): that partner must be pinned by a Dirichlet or Periodic row,
    // otherwise the row couples two free unknowns (checked on the host so that every rank takes the same decision)
    for (int lo = 0; lo < 6; lo += 2) {
        const int hi = lo + 1;
        if (s->bd.kind[hi] == PB200_BC_PERIODIC && s->bd.present[lo] && s->bd.kind[lo] != PB200_BC_DIRICHLET && s->bd.kind[lo] != PB200_BC_PERIODIC)
            return set_err(ctx, PB200_EUNSUPPORTED, "a Periodic border row couples to a FREE unknown (the opposite side has neither a Dirichlet nor a Periodic row): "
                                                    "such coupling rows are not supported; Periodic on both sides or Periodic + Dirichlet is");
    }
    int *d_err = nullptr, h_err = 0;
    CUDA_TRY(ctx, cudaMalloc((void **)&d_err, sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    DISPATCH_N(g.N, (k_build_masks<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->p2, ct1, ct2, s->sp, s->bd, s->m1, s->m2, s->ufix1, s->ufix2, d_err)));
    LAUNCH_CHECK(ctx);
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_err);
    (void)h_err;   // cannot trigger after the host check above (kept as a device-side assertion)
    double *fl[2] = {s->ufix1, s->ufix2};
    int rc;
    if ((rc = halo_exchange(ctx, g, fl, 2))) return rc;
    k_count<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m1, s->sp.phase_type == PB200_DIPH ? s->m2 : nullptr, ctx->d_partials, ctx->d_results + SL_TMP,
                                                   ctx->d_counter);
    LAUNCH_CHECK(ctx);
    if ((rc = allreduce_results(ctx, SL_TMP, 2))) return rc;
    double cnt[2];
    if ((rc = fetch_results(ctx, SL_TMP, 2, cnt))) return rc;
    s->dof_bulk = (int64_t)(cnt[0] + 0.5);
    s->dof_ifc = (int64_t)(cnt[1] + 0.5);
    {
        if (s->Kcell) { cudaFree(s->Kcell); s->Kcell = nullptr; }
        const unsigned char *ma = s->m1, *mb = s->sp.phase_type == PB200_DIPH ? s->m2 : nullptr;
        Grid gg = g;
        cudaStream_t st = ctx->stream;
        if ((rc = fold_compact(ctx, [&](long long *list, int *cnt, int cap) { k_mark_known_rows<<<grid, RED_THREADS, 0, st>>>(gg, ma, mb, list, cnt, cap); }, &s->Kcell, &s->nK))) return rc;
    }
    // the solution vector is only written where unknowns are active: clear what an earlier mask set may have left elsewhere
    for (int f = 0; f < s->nf; ++f) CUDA_TRY(ctx, cudaMemsetAsync(s->x.f[f], 0, sizeof(double) * (size_t)g.nloc, ctx->stream));
    s->masks_dirty = false;
    s->diag_key.cV = -1;
    s->F.built = false;
    return PB200_OK;
}

static int stage_src(pb200_solver *s, double **slot, const double *host, double cst, SrcSpec *out)
{
    out->cst = cst; out->arr = nullptr;
    if (!host) return PB200_OK;
    int rc;
    if (!*slot && (rc = dev_alloc(s->ctx, slot, s->g.nloc))) return rc;
    if ((rc = upload_owned(s->ctx, s->g, *slot, host))) return rc;
    out->arr = *slot;
    return PB200_OK;
}

extern "C" int pb200_solver_step(pb200_solver *s, const pb200_step_in *in, const pb200_krylov_opts *opts_in, pb200_step_stats *stats)
{
    if (!s || !in) return set_err(nullptr, PB200_EINVAL, "NULL argument");
    if (!s->parts.empty()) return team_solver_step(s, in, opts_in, stats);
    pb200_ctx *ctx = s->ctx;
    const Grid &g = s->g;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches, applies0 = ctx->apply_launches;
    ctx->pev_used = 0;
    pb200_krylov_opts o = {PB200_KRYLOV_AUTO, 1e-10, 0.0, 10000, 1, 1, PB200_PATH_AUTO};
    if (opts_in) o = *opts_in;
    if (o.maxit <= 0) o.maxit = 10000;
    if (o.check_every <= 0) o.check_every = 1;
    const bool diph = s->sp.phase_type == PB200_DIPH;
    const bool unsteady = s->sp.time_type == PB200_UNSTEADY;
    const bool cn = unsteady && in->scheme == PB200_CN;
    int method = o.method;
    const bool nonconstD = s->D1arr != nullptr;
    if (o.path == PB200_PATH_FOLDED && !fold_eligible(s))
        return set_err(ctx, PB200_EUNSUPPORTED, "the folded path needs jump / Robin coefficients of one sign (alpha2/alpha1, beta1, beta2 > 0; beta > 0, alpha >= 0)");
    const bool use_fold = o.path != PB200_PATH_GENERIC && fold_eligible(s);
    if (o.precond != PB200_PRECOND_DEFAULT && o.precond != PB200_PRECOND_MG) return set_err(ctx, PB200_EINVAL, "unknown preconditioner (pb200_krylov_opts.precond)");
    if (o.precond == PB200_PRECOND_MG && !use_fold)   // never a silent fall-back to the plain iteration
        return set_err(ctx, PB200_EUNSUPPORTED, "the multigrid preconditioner belongs to the folded path (PB200_PATH_AUTO / PB200_PATH_FOLDED on an eligible system)");
    // AUTO: the folded system is symmetric positive definite for mono AND diphasic problems, so CG (one operator apply per
    // iteration) is the cheaper choice there; the reference's rows of the diphasic system are not symmetric => BiCGSTAB.
    const bool advect = s->p1.kd || s->p2.kd;   // ConvectionOps: non-symmetric rows (gmres in the reference, src/solver/advectiondiffusion.jl:61) -> BiCGSTAB
    if (method == PB200_KRYLOV_AUTO) method = ((use_fold || !diph) && !advect) ? PB200_KRYLOV_CG : PB200_KRYLOV_BICGSTAB;
    if (method == PB200_KRYLOV_CG && advect) return set_err(ctx, PB200_EUNSUPPORTED, "CG on an advection-diffusion system (not symmetric); use BiCGSTAB");
    if (advect && diph && in->scheme == PB200_CN)   // the reference's diphasic CN right-hand side drops the explicit diffusion part (advectiondiffusion.jl:372-376): not reproduced
        return set_err(ctx, PB200_EUNSUPPORTED, "diphasic advection-diffusion with Crank-Nicolson is not supported (BE is)");
    if (method == PB200_KRYLOV_CG && diph && !use_fold)
        return set_err(ctx, PB200_EUNSUPPORTED, "CG on the diphasic system needs the folded (symmetrised) path; use BiCGSTAB");
    (void)nonconstD;
    int rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (s->masks_dirty) { if ((rc = build_masks(s))) return rc; }
    else if (s->values_dirty) {
        k_refresh_ufix<<<sgrid(ctx, g.nown), RED_THREADS, 0, ctx->stream>>>(g, s->bd, s->m1, s->sp.phase_type == PB200_DIPH ? s->m2 : nullptr, s->ufix1, s->ufix2);
        LAUNCH_CHECK(ctx);
        double *fl[2] = {s->ufix1, s->ufix2};
        if ((rc = halo_exchange(ctx, g, fl, 2))) return rc;
    }
    s->values_dirty = false;

    StepCoef sc;
    if (unsteady) {
        if (!(in->dt > 0.0)) return set_err(ctx, PB200_EINVAL, "dt must be positive");
        sc.cV = 1.0; sc.cn = cn ? 1 : 0;
        sc.c = cn ? 0.5 * in->dt : in->dt; sc.ce = cn ? 0.5 * in->dt : 0.0; sc.c2 = cn ? 0.5 * in->dt : 1.0;
        sc.wf0 = cn ? 0.5 * in->dt : 0.0; sc.wf1 = cn ? 0.5 * in->dt : in->dt;
        sc.wg0 = cn ? 1.0 : 0.0; sc.wg1 = 1.0;
    } else {
        sc.cV = 0.0; sc.cn = 0; sc.c = 1.0; sc.ce = 0.0; sc.c2 = 1.0; sc.wf0 = 1.0; sc.wf1 = 0.0; sc.wg0 = 1.0; sc.wg1 = 0.0;
    }
    sc.sym = (!use_fold && method == PB200_KRYLOV_CG) ? 1 : 0;
    ApplyCoef ac = {sc.cV, sc.c, sc.c2, sc.sym, 1};

    // sources
    SrcSpec f[2][2], gsp[2];
    for (int ph = 0; ph < 2; ++ph)
        for (int w = 0; w < 2; ++w)
            if ((rc = stage_src(s, &s->fS[ph][w], in->f_arr[ph][w], in->f_const[ph][w], &f[ph][w]))) return rc;
    for (int w = 0; w < 2; ++w)
        if ((rc = stage_src(s, &s->gS[w], in->g_arr[w], in->g_const[w], &gsp[w]))) return rc;

    const int grid = sgrid(ctx, g.nown);
    // initial guess: 0 none, 1 previous state, 2 linear extrapolation from the two previous states (needs one earlier step with the same dt)
    int warm = (!unsteady || !o.warm_start) ? 0 : 1;
    if (warm && o.warm_start >= 2 && in->dt == s->prev_dt) warm = 1 + (s->n_prev < o.warm_start - 1 ? s->n_prev : o.warm_start - 1);
    if (warm > PB_MAXHIST) warm = PB_MAXHIST;
    auto guess_spec = [&](double *cur, const std::vector<double *> &hist) {
        GuessSpec gsx;
        gsx.m = warm;
        double binom = 1.0;   // c_j = (-1)^j C(m, j+1)
        for (int j = 0; j < PB_MAXHIST; ++j) {
            gsx.T[j] = nullptr; gsx.c[j] = 0.0;
            if (j < warm) {
                binom = binom * (double)(warm - j) / (double)(j + 1);
                gsx.c[j] = (j & 1) ? -binom : binom;
                gsx.T[j] = j == 0 ? cur : hist[j - 1];
            }
        }
        return gsx;
    };
    // folded system (cached per coefficient set and mask set)
    if (use_fold && (!s->F.built || s->F.key[0] != ac.cV || s->F.key[1] != ac.c || s->F.key[2] != ac.c2)) { if ((rc = fold_build(s, ac)) || (rc = fold_poly_setup(s))) return rc; }
    // steps without an explicit operator part on the folded path: b, b^ and the scaled initial guess come out of ONE pass over the active
    // tiles (kf_rhs_dense) instead of k_rhs_* + kf_to_scaled_dense + kf_guess_dense over all cells
    const bool fused_rhs = use_fold && !getenv("PB200_NO_FUSED_RHS") && (!cn || (sc.ce == sc.c && !getenv("PB200_NO_FUSED_CN")));
    GuessSpec fgs[3];
    if (use_fold) {
        fgs[0] = guess_spec(s->Tw[0], s->histW[0]);
        fgs[1] = diph ? guess_spec(s->Tw[1], s->histW[1]) : fgs[0];
        fgs[2] = diph ? guess_spec(s->Tg[1], s->histG[1]) : guess_spec(s->Tg[0], s->histG[0]);
    }
    auto rhs_fused = [&]() -> int {
        FoldSys &F = s->F;
        RhsSrc S;
        for (int ph = 0; ph < 2; ++ph) { S.f0[ph] = f[ph][0]; S.f1[ph] = f[ph][1]; S.Tw[ph] = s->Tw[ph]; }
        S.m[0] = s->m1; S.m[1] = diph ? s->m2 : s->m1;
        FVec vexp; for (int q = 0; q < 3; ++q) vexp.f[q] = nullptr;
        if (cn) {
            // Crank-Nicolson: explicit part = M^ x^n with x^n = L^T T^n (kf_rhs_dense); F.p and F.v are free until the solve starts
            auto cur = [&](double *T) { GuessSpec q; q.m = 1; for (int j = 0; j < PB_MAXHIST; ++j) { q.T[j] = nullptr; q.c[j] = 0.0; } q.T[0] = T; q.c[0] = 1.0; return q; };
            const GuessSpec c0 = cur(s->Tw[0]), c1 = diph ? cur(s->Tw[1]) : c0, cw = diph ? cur(s->Tg[1]) : cur(s->Tg[0]);
            const bool band = F.d.has_w && F.d.nBown > 0;
            kf_guess_dense<<<wave_grid(s, kf_guess_dense), FCH, 0, ctx->stream>>>(F.d, F.I, c0, c1, F.p); LAUNCH_CHECK(ctx);
            if (band) { kf_guess_band<<<band_grid(F.d.nBown), 128, 0, ctx->stream>>>(F.d, c0, c1, cw, F.p); LAUNCH_CHECK(ctx); }
            int rc2;
            if ((rc2 = fold_apply(s, F.p, F.v, F.v, 0))) return rc2;
            vexp = F.v;
        }
        kf_rhs_dense<<<wave_grid(s, kf_rhs_dense), FCH, 0, ctx->stream>>>(F.d, F.I, sc, S, fgs[0], fgs[1], s->b, F.b, F.x, vexp); LAUNCH_CHECK(ctx);
        return PB200_OK;
    };
    // ghost planes of the state (stencil inputs of the explicit part)
    {
        double *fl[4] = {s->Tw[0], s->Tg[0], s->Tw[1], s->Tg[1]};
        if (cn && (rc = halo_exchange(ctx, g, fl, diph ? 4 : 2))) return rc;
    }
    if (!diph) {
        k_gamma_known<<<grid, RED_THREADS, 0, ctx->stream>>>(g, sc, s->m1, gsp[0], gsp[1], s->Tg[0], s->gK);
        LAUNCH_CHECK(ctx);
        double *fl[1] = {s->gK};
        if ((rc = halo_exchange(ctx, g, fl, 1))) return rc;
        if (fused_rhs) { if ((rc = rhs_fused())) return rc; }
        else {
            DISPATCH_N(g.N, (k_rhs_mono<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->sp, sc, s->m1, s->Tw[0], s->Tg[0], s->ufix1, s->gK, f[0][0],
                                                                                   f[0][1], gsp[0], gsp[1], s->b.f[0], s->nf == 2 ? s->b.f[1] : nullptr, 1)));
            LAUNCH_CHECK(ctx);
        }
        if (fused_rhs && cn) {   // the listed rows in full: unfolded explicit part + known part
            if (s->nK > 0) {
                DISPATCH_N(g.N, (k_rhs_mono<N><<<(s->nK + 127) / 128, 128, 0, ctx->stream>>>(g, s->p1, s->sp, sc, s->m1, s->Tw[0], s->Tg[0], s->ufix1, s->gK, f[0][0],
                                                                                        f[0][1], gsp[0], gsp[1], s->b.f[0], s->nf == 2 ? s->b.f[1] : nullptr, 0,
                                                                                        s->Kcell, s->nK)));
                LAUNCH_CHECK(ctx);
            }
        } else if (s->nK > 0) {
            DISPATCH_N(g.N, (k_rhs_known_mono<N><<<(s->nK + 127) / 128, 128, 0, ctx->stream>>>(g, s->p1, s->sp, sc, s->m1, s->Kcell, s->nK, s->ufix1, s->gK, s->b.f[0],
                                                                                              s->nf == 2 ? s->b.f[1] : nullptr, fused_rhs ? 1 : 0, gsp[0], gsp[1])));
            LAUNCH_CHECK(ctx);
        }
        if (!use_fold) {   // (the folded path evaluates the guess itself, in scaled unknowns)
            k_guess<<<grid, RED_THREADS, 0, ctx->stream>>>(g, guess_spec(s->Tw[0], s->histW[0]), s->m1, MB_FREE, s->x.f[0]);
            LAUNCH_CHECK(ctx);
            if (s->nf == 2) { k_guess<<<grid, RED_THREADS, 0, ctx->stream>>>(g, guess_spec(s->Tg[0], s->histG[0]), s->m1, MB_IFREE, s->x.f[1]); LAUNCH_CHECK(ctx); }
        }
    } else {
        if (gsp[0].arr) { double *fl[1] = {s->gS[0]}; if ((rc = halo_exchange(ctx, g, fl, 1))) return rc; }
        if (fused_rhs) { if ((rc = rhs_fused())) return rc; }
        else {
            DISPATCH_N(g.N, (k_rhs_diph<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->p2, s->sp, sc, s->m1, s->m2, s->Tw[0], s->Tg[0], s->Tw[1],
                                                                                   s->Tg[1], s->ufix1, s->ufix2, f[0][0], f[0][1], f[1][0], f[1][1], gsp[0],
                                                                                   gsp[1], s->b.f[0], s->b.f[1], s->b.f[2], 1)));
            LAUNCH_CHECK(ctx);
        }
        if (fused_rhs && cn) {   // the listed rows in full: unfolded explicit part + known part
            if (s->nK > 0) {
                DISPATCH_N(g.N, (k_rhs_diph<N><<<(s->nK + 127) / 128, 128, 0, ctx->stream>>>(g, s->p1, s->p2, s->sp, sc, s->m1, s->m2, s->Tw[0], s->Tg[0], s->Tw[1],
                                                                                        s->Tg[1], s->ufix1, s->ufix2, f[0][0], f[0][1], f[1][0], f[1][1], gsp[0],
                                                                                        gsp[1], s->b.f[0], s->b.f[1], s->b.f[2], 0, s->Kcell, s->nK)));
                LAUNCH_CHECK(ctx);
            }
        } else if (s->nK > 0) {
            DISPATCH_N(g.N, (k_rhs_known_diph<N><<<(s->nK + 127) / 128, 128, 0, ctx->stream>>>(g, s->p1, s->p2, s->sp, sc, s->m1, s->m2, s->Kcell, s->nK, s->ufix1, s->ufix2,
                                                                                              gsp[0], s->b.f[0], s->b.f[1], s->b.f[2], fused_rhs ? 1 : 0, gsp[1])));
            LAUNCH_CHECK(ctx);
        }
        if (!use_fold) {
            k_guess<<<grid, RED_THREADS, 0, ctx->stream>>>(g, guess_spec(s->Tw[0], s->histW[0]), s->m1, MB_FREE, s->x.f[0]); LAUNCH_CHECK(ctx);
            k_guess<<<grid, RED_THREADS, 0, ctx->stream>>>(g, guess_spec(s->Tw[1], s->histW[1]), s->m2, MB_FREE, s->x.f[1]); LAUNCH_CHECK(ctx);
            k_guess<<<grid, RED_THREADS, 0, ctx->stream>>>(g, guess_spec(s->Tg[1], s->histG[1]), s->m2, MB_IFREE, s->x.f[2]); LAUNCH_CHECK(ctx);
        }
    }
    if (!use_fold && !s->have_generic) {
        MVec *vs[] = {&s->r, &s->r0, &s->p, &s->ph, &s->v, &s->s, &s->sh, &s->t, &s->dinv};
        for (MVec *v : vs) if ((rc = solver_vec(s, v))) return rc;
        s->have_generic = true;
    }
    if (use_fold) {
    } else if (memcmp(&ac, &s->diag_key, sizeof(ac)) != 0) {   // Jacobi diagonal (cached per coefficient set)
        if (!diph) {
            DISPATCH_N(g.N, (k_diag_mono<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->sp, ac, s->m1, s->dinv.f[0], s->nf == 2 ? s->dinv.f[1] : nullptr)));
        } else {
            DISPATCH_N(g.N, (k_diag_diph<N><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->p1, s->p2, s->sp, ac, s->m1, s->m2, s->dinv.f[0], s->dinv.f[1], s->dinv.f[2])));
        }
        LAUNCH_CHECK(ctx);
        k_recip<<<dim3(grid, s->nf), RED_THREADS, 0, ctx->stream>>>(g, s->dinv);
        LAUNCH_CHECK(ctx);
        s->diag_key = ac;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));

    // ---- Krylov ----------------------------------------------------------------------------------------------------------
    double *res = ctx->d_results;
    const dim3 vgrid(grid, s->nf);
    MVec z = s->x;  // alias
    int it = 0, converged = 0;
    double rnorm = 0.0, bnorm = 0.0;
    if (use_fold) {
        if ((rc = fold_solve(s, method, o, fgs, fused_rhs, &it, &converged, &rnorm, &bnorm))) return rc;
    } else {
    // bnorm
    k_dots<1><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->b, s->b, s->b, s->b, ctx->d_partials, res + SL_BB, ctx->d_counter); LAUNCH_CHECK(ctx);
    if ((rc = allreduce_results(ctx, SL_BB, 1))) return rc;
    // r = b - A x
    if ((rc = apply_op(s, ac, z, s->t))) return rc;
    int cur = 0;  // pair index holding (rho, rr)
    k_resid<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->b, s->t, s->r, ctx->d_partials, res + SL_TMP, ctx->d_counter); LAUNCH_CHECK(ctx);
    if ((rc = allreduce_results(ctx, SL_TMP, 1))) return rc;
    double h[2];
    if ((rc = fetch_results(ctx, SL_BB, 2, h))) return rc;   // SL_BB, SL_TMP adjacent
    bnorm = sqrt(h[0]);
    rnorm = sqrt(h[1]);
    const double tol = fmax(o.rtol * bnorm, o.atol);
    converged = rnorm <= tol ? 1 : 0;
    if (!converged) {
        if (method == PB200_KRYLOV_CG) {
            // p = dinv r ; rho = (r, dinv r)
            k_scale<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, s->dinv, s->r, s->p); LAUNCH_CHECK(ctx);
            k_dots<1><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->r, s->p, s->r, s->p, ctx->d_partials, res + 2 * cur, ctx->d_counter); LAUNCH_CHECK(ctx);
            if ((rc = allreduce_results(ctx, 2 * cur, 1))) return rc;
            while (it < o.maxit) {
                if ((rc = apply_op(s, ac, s->p, s->v))) return rc;
                k_dots<1><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->p, s->v, s->p, s->v, ctx->d_partials, res + SL_SIGMA, ctx->d_counter); LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, SL_SIGMA, 1))) return rc;
                const int nxt = cur ^ 1;
                k_cg_update<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, res, 2 * cur, 2 * nxt, s->dinv, s->p, s->v, z, s->r, ctx->d_partials, ctx->d_counter); LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, 2 * nxt, 2))) return rc;
                k_cg_p<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, res, 2 * cur, 2 * nxt, s->dinv, s->r, s->p); LAUNCH_CHECK(ctx);
                cur = nxt;
                ++it;
                if (it % o.check_every == 0 || it == o.maxit) {
                    if ((rc = fetch_results(ctx, 2 * cur + 1, 1, h))) return rc;
                    rnorm = sqrt(h[0]);
                    if (rnorm <= tol) { converged = 1; break; }
                    if (!(rnorm == rnorm)) break;
                }
            }
        } else {
            // r0 = r ; p = r ; ph = dinv p ; rho = (r0, r) = rr
            k_copy<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, s->r, s->r0); LAUNCH_CHECK(ctx);
            k_copy<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, s->r, s->p); LAUNCH_CHECK(ctx);
            k_scale<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, s->dinv, s->p, s->ph); LAUNCH_CHECK(ctx);
            k_dots<2><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->r0, s->r, s->r, s->r, ctx->d_partials, res + 2 * cur, ctx->d_counter); LAUNCH_CHECK(ctx);
            if ((rc = allreduce_results(ctx, 2 * cur, 2))) return rc;
            while (it < o.maxit) {
                if ((rc = apply_op(s, ac, s->ph, s->v))) return rc;
                k_dots<1><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->r0, s->v, s->r0, s->v, ctx->d_partials, res + SL_SIGMA, ctx->d_counter); LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, SL_SIGMA, 1))) return rc;
                k_bicg_s<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, res, 2 * cur, s->dinv, s->r, s->v, s->s, s->sh); LAUNCH_CHECK(ctx);
                if ((rc = apply_op(s, ac, s->sh, s->t))) return rc;
                k_dots<2><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, s->t, s->s, s->t, s->t, ctx->d_partials, res + SL_TS, ctx->d_counter); LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, SL_TS, 2))) return rc;
                const int nxt = cur ^ 1;
                k_bicg_xr<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->nf, res, 2 * cur, 2 * nxt, s->ph, s->sh, s->s, s->t, s->r0, z, s->r, ctx->d_partials, ctx->d_counter); LAUNCH_CHECK(ctx);
                if ((rc = allreduce_results(ctx, 2 * nxt, 2))) return rc;
                k_bicg_p<<<vgrid, RED_THREADS, 0, ctx->stream>>>(g, res, 2 * cur, 2 * nxt, s->dinv, s->r, s->v, s->p, s->ph); LAUNCH_CHECK(ctx);
                cur = nxt;
                ++it;
                if (it % o.check_every == 0 || it == o.maxit) {
                    if ((rc = fetch_results(ctx, 2 * cur + 1, 1, h))) return rc;
                    rnorm = sqrt(h[0]);
                    if (rnorm <= tol) { converged = 1; break; }
                    if (!(rnorm == rnorm)) break;
                }
            }
        }
    }
    }   // generic path
    // ---- write the new state ----------------------------------------------------------------------------------------------
    if (s->copy_pending) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, s->copy_done, 0));   // pb200_solver_get_state_async still reads the old one
    if (unsteady && o.warm_start >= 2) {
        // keep the last states for the next step's extrapolated guess: the new state is written into the oldest buffer (rotation, no copy)
        const int np = diph ? 2 : 1;
        int keep = o.warm_start - 1;
        if (keep > PB_MAXHIST - 1) keep = PB_MAXHIST - 1;
        for (int ph = 0; ph < np; ++ph) {
            std::vector<double *> *hs[2] = {&s->histW[ph], &s->histG[ph]};
            double **cur[2] = {&s->Tw[ph], &s->Tg[ph]};
            for (int q = 0; q < 2; ++q) {
                std::vector<double *> &h = *hs[q];
                double *fresh = nullptr;
                if ((int)h.size() < keep) { if ((rc = dev_alloc(ctx, &fresh, g.nloc))) return rc; }
                else { fresh = h.back(); h.pop_back(); }
                h.insert(h.begin(), *cur[q]);     // T^n becomes T^(n-1)
                *cur[q] = fresh;                  // receives T^(n+1)
            }
        }
        s->n_prev = s->prev_dt == in->dt ? (s->n_prev < keep ? s->n_prev + 1 : keep) : 1;
        s->prev_dt = in->dt;
    } else s->n_prev = 0;
    if (has_slave_rows(s)) {   // x_row = x_adj (+ g dx from ufix in k_store_bulk)
        k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m1, z.f[0], 0); LAUNCH_CHECK(ctx);
        if (diph) { k_slave_copy<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->m2, z.f[1], 0); LAUNCH_CHECK(ctx); }
    }
    k_store_bulk<<<grid, RED_THREADS, 0, ctx->stream>>>(g, z.f[0], s->ufix1, s->Tw[0]); LAUNCH_CHECK(ctx);
    if (!diph) {
        const double *src = s->nf == 2 ? z.f[1] : s->gK;
        CUDA_TRY(ctx, cudaMemcpyAsync(s->Tg[0] + g.plane, src + g.plane, sizeof(double) * (size_t)g.nown, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        k_store_bulk<<<grid, RED_THREADS, 0, ctx->stream>>>(g, z.f[1], s->ufix2, s->Tw[1]); LAUNCH_CHECK(ctx);
        k_store_diph_ifc<<<grid, RED_THREADS, 0, ctx->stream>>>(g, s->sp, gsp[0], z.f[2], s->Tg[0], s->Tg[1]); LAUNCH_CHECK(ctx);
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev2));
    if (stats) {
        float ms_setup = 0, ms_solve = 0;
        cudaEventElapsedTime(&ms_setup, ctx->ev0, ctx->ev1);
        cudaEventElapsedTime(&ms_solve, ctx->ev1, ctx->ev2);
        stats->iters = it; stats->converged = converged; stats->rnorm = rnorm; stats->bnorm = bnorm;
        stats->solve_ms = ms_solve; stats->setup_ms = ms_setup;
        stats->dof_bulk = s->dof_bulk; stats->dof_ifc = s->dof_ifc;
        stats->launches = ctx->launches - launches0;
        prof_collect(ctx, stats->kernel_ms, stats->kernel_launches);
        stats->apply_ms = stats->kernel_ms[PB_PROF_APPLY];
        stats->apply_launches = ctx->apply_launches - applies0;
        stats->apply_cells_uniform = use_fold ? s->F.cells_uniform : 0;
        stats->apply_cells_general = use_fold ? s->F.cells_general : 0;
        stats->apply_cells_fast = use_fold ? s->F.cells_fast : 0;
        stats->band_cells = use_fold ? s->F.d.nBown : 0;
        stats->band_rows = use_fold ? s->F.d.nE : 0;
    }
    if (!converged) return set_err(ctx, PB200_ENOTCONV, "Krylov solve did not reach the tolerance within maxit iterations");
    return PB200_OK;
}


// =================================================================================================================================
// One process driving several GPUs -- pb200_init_multi ("Julia is one process"; the reference's scripts are single-process).
// The team handle stands for ndev member contexts, rank r on devices[r]; every entry point fans out to the members on one host thread per
// GPU (the calls contain collective steps -- halo exchanges, reductions -- that all ranks must reach together) and cuts / reassembles the
// host arrays: per-cell arrays keep the reference's GLOBAL padded length n = prod(n_i + 1); member r owns the planes [k0_r, k1_r) of the
// slowest dimension, a contiguous range of every block of n doubles.
// =================================================================================================================================
template <typename F>
static int team_run(pb200_ctx *tc, F f)
{
    Team *T = tc->team;
    std::vector<int> rc(T->n, 0);
    std::vector<std::string> err(T->n);
    std::vector<std::thread> th;
    for (int r = 0; r < T->n; ++r)
        th.emplace_back([&, r]() {
            if (T->ctx[r]) cudaSetDevice(T->ctx[r]->device);   // (the members do not exist yet while pb200_init_multi creates them)
            rc[r] = f(r);
            if (rc[r]) err[r] = g_last_error;
        });
    for (std::thread &t : th) t.join();
    for (int r = 0; r < T->n; ++r)
        if (rc[r]) return set_err(tc, rc[r], "rank " + std::to_string(r) + ": " + err[r]);
    return PB200_OK;
}

extern "C" int pb200_init_multi(pb200_ctx **ctx, const int *devices, int ndev)
{
    if (!ctx || !devices || ndev < 1 || ndev > P2P_MAXR) return set_err(nullptr, PB200_EINVAL, "need 1 .. 8 devices");
    if (ndev == 1) return pb200_init(ctx, devices[0]);
    char id[128];
    int rc = pb200_nccl_unique_id(id);
    if (rc) return rc;
    pb200_ctx *tc = new pb200_ctx();
    Team *T = new Team();
    T->n = ndev; T->ctx.assign(ndev, nullptr); T->mbox.assign(ndev, nullptr);
    tc->team = T; tc->is_team = true; tc->device = devices[0]; tc->nranks = ndev;
    rc = team_run(tc, [&](int r) { return pb200_init_dist(&T->ctx[r], devices[r], r, ndev, id); });   // ncclCommInitRank: all ranks together
    if (rc) {
        for (pb200_ctx *m : T->ctx) if (m) pb200_finalize(m);
        const std::string msg = tc->err;
        delete T; delete tc;
        return set_err(nullptr, rc, msg);
    }
    for (pb200_ctx *m : T->ctx) m->team = T;
    *ctx = tc;
    return PB200_OK;
}

// blocks of `n` doubles of a global host array <-> blocks of nown doubles of member r
static void team_cut(const Grid &gr, int64_t ntot, const double *glob, int nblk, std::vector<double> &loc)
{
    loc.resize((size_t)nblk * gr.nown);
    for (int b = 0; b < nblk; ++b) memcpy(loc.data() + (size_t)b * gr.nown, glob + (size_t)b * ntot + (size_t)gr.k0 * gr.plane, sizeof(double) * (size_t)gr.nown);
}
static void team_paste(const Grid &gr, int64_t ntot, const std::vector<double> &loc, int nblk, double *glob)
{
    for (int b = 0; b < nblk; ++b) memcpy(glob + (size_t)b * ntot + (size_t)gr.k0 * gr.plane, loc.data() + (size_t)b * gr.nown, sizeof(double) * (size_t)gr.nown);
}
static int team_global_grid(pb200_ctx *tc, int ndim, const int *n, const double *x0, const double *L, Grid *g)
{
    pb200_ctx one;          // (a single-rank description of the whole grid: sizes only)
    one.rank = 0; one.nranks = 1;
    return make_grid(&one, ndim, n, x0, L, g) ? set_err(tc, PB200_EINVAL, g_last_error) : PB200_OK;
}

static int team_capacity_create(pb200_ctx *tc, int ndim, const int *n, const double *x0, const double *L, const pb200_levelset *ls, int cc, pb200_capacity **out)
{
    Team *T = tc->team;
    pb200_capacity *c = new pb200_capacity();
    c->ctx = tc; c->has_cg = cc != 0;
    int rc = team_global_grid(tc, ndim, n, x0, L, &c->g);
    if (rc) { delete c; return rc; }
    c->parts.assign(T->n, nullptr);
    rc = team_run(tc, [&](int r) { return pb200_capacity_create(T->ctx[r], ndim, n, x0, L, ls, cc, &c->parts[r]); });
    if (rc) { pb200_capacity_destroy(c); return rc; }
    *out = c;
    return PB200_OK;
}
static int team_capacity_import(pb200_ctx *tc, int ndim, const int *n, const double *x0, const double *L, const double *V, const double *Gamma, const double *ct,
                                const double *A, const double *B, const double *W, const double *Co, const double *Cg, pb200_capacity **out)
{
    Team *T = tc->team;
    pb200_capacity *c = new pb200_capacity();
    c->ctx = tc; c->has_cg = Cg != nullptr;
    int rc = team_global_grid(tc, ndim, n, x0, L, &c->g);
    if (rc) { delete c; return rc; }
    c->parts.assign(T->n, nullptr);
    const int64_t nt = c->g.ntot;
    rc = team_run(tc, [&](int r) {
        Grid gr;
        int rr = make_grid(T->ctx[r], ndim, n, x0, L, &gr);
        if (rr) return rr;
        std::vector<double> a, b, w, co, cg;
        if (A) team_cut(gr, nt, A, ndim, a);
        if (B) team_cut(gr, nt, B, ndim, b);
        if (W) team_cut(gr, nt, W, ndim, w);
        if (Co) team_cut(gr, nt, Co, ndim, co);
        if (Cg) team_cut(gr, nt, Cg, ndim, cg);
        const size_t off = (size_t)gr.k0 * gr.plane;     // single blocks: a pointer offset is enough
        return pb200_capacity_import(T->ctx[r], ndim, n, x0, L, V ? V + off : nullptr, Gamma ? Gamma + off : nullptr, ct ? ct + off : nullptr, A ? a.data() : nullptr,
                                     B ? b.data() : nullptr, W ? w.data() : nullptr, Co ? co.data() : nullptr, Cg ? cg.data() : nullptr, &c->parts[r]);
    });
    if (rc) { pb200_capacity_destroy(c); return rc; }
    *out = c;
    return PB200_OK;
}
static int team_capacity_export(pb200_capacity *c, double *V, double *Gamma, double *ct, double *A, double *B, double *W, double *Co, double *Cg)
{
    pb200_ctx *tc = c->ctx;
    const int64_t nt = c->g.ntot;
    const int N = c->g.N;
    return team_run(tc, [&](int r) {
        const Grid &gr = c->parts[r]->g;
        std::vector<double> a(A ? (size_t)N * gr.nown : 0), b(B ? (size_t)N * gr.nown : 0), w(W ? (size_t)N * gr.nown : 0), co(Co ? (size_t)N * gr.nown : 0),
            cg(Cg ? (size_t)N * gr.nown : 0);
        const size_t off = (size_t)gr.k0 * gr.plane;
        int rr = pb200_capacity_export(c->parts[r], V ? V + off : nullptr, Gamma ? Gamma + off : nullptr, ct ? ct + off : nullptr, A ? a.data() : nullptr,
                                       B ? b.data() : nullptr, W ? w.data() : nullptr, Co ? co.data() : nullptr, Cg ? cg.data() : nullptr);
        if (rr) return rr;
        if (A) team_paste(gr, nt, a, N, A);
        if (B) team_paste(gr, nt, b, N, B);
        if (W) team_paste(gr, nt, w, N, W);
        if (Co) team_paste(gr, nt, co, N, Co);
        if (Cg) team_paste(gr, nt, cg, N, Cg);
        return PB200_OK;
    });
}
static int team_ops_create(pb200_capacity *cap, pb200_ops **out)
{
    pb200_ctx *tc = cap->ctx;
    pb200_ops *o = new pb200_ops();
    o->ctx = tc; o->cap = cap;
    o->parts.assign(tc->team->n, nullptr);
    int rc = team_run(tc, [&](int r) { return pb200_ops_create(cap->parts[r], &o->parts[r]); });
    if (rc) { pb200_ops_destroy(o); return rc; }
    *out = o;
    return PB200_OK;
}
// what: 0 export W! (out: N blocks), 1 grad (a: 2 blocks, out: N blocks), 2 div (a, b: N blocks each, out: 1 block)
static int team_ops_vec(pb200_ops *o, int what, const double *a, const double *b, double *out)
{
    pb200_ctx *tc = o->ctx;
    const int64_t nt = o->cap->g.ntot;
    const int N = o->cap->g.N;
    return team_run(tc, [&](int r) {
        const Grid &gr = o->parts[r]->cap->g;
        std::vector<double> la, lb, lo((size_t)(what == 2 ? 1 : N) * gr.nown);
        int rr;
        if (what == 0) rr = pb200_ops_export_wdag(o->parts[r], lo.data());
        else if (what == 1) { team_cut(gr, nt, a, 2, la); rr = pb200_ops_grad(o->parts[r], la.data(), lo.data()); }
        else { team_cut(gr, nt, a, N, la); team_cut(gr, nt, b, N, lb); rr = pb200_ops_div(o->parts[r], la.data(), lb.data(), lo.data()); }
        if (rr) return rr;
        team_paste(gr, nt, lo, what == 2 ? 1 : N, out);
        return PB200_OK;
    });
}
static int team_solver_create(pb200_ctx *tc, const pb200_solver_desc *d, pb200_solver **out)
{
    Team *T = tc->team;
    if (d->ops1->parts.empty() || (d->ops2 && d->ops2->parts.empty())) return set_err(tc, PB200_EINVAL, "the operators do not belong to this team context");
    pb200_solver *s = new pb200_solver();
    s->ctx = tc; s->g = d->ops1->cap->g;
    s->team_nblk = d->phase_type == PB200_DIPH ? 4 : 2;
    s->parts.assign(T->n, nullptr);
    int rc = team_run(tc, [&](int r) {
        pb200_solver_desc dr = *d;
        dr.ops1 = d->ops1->parts[r];
        dr.ops2 = d->ops2 ? d->ops2->parts[r] : nullptr;
        const Grid &gr = dr.ops1->cap->g;
        const size_t off = (size_t)gr.k0 * gr.plane;
        if (d->D1_arr) dr.D1_arr = d->D1_arr + off;
        if (d->D2_arr) dr.D2_arr = d->D2_arr + off;
        return pb200_solver_create(T->ctx[r], &dr, &s->parts[r]);
    });
    if (rc) { pb200_solver_destroy(s); return rc; }
    *out = s;
    return PB200_OK;
}
static int team_solver_state(pb200_solver *s, const double *in, double *out)
{
    pb200_ctx *tc = s->ctx;
    const int64_t nt = s->g.ntot;
    return team_run(tc, [&](int r) {
        const Grid &gr = s->parts[r]->g;
        std::vector<double> loc;
        if (in) { team_cut(gr, nt, in, s->team_nblk, loc); return pb200_solver_set_state(s->parts[r], loc.data()); }
        loc.resize((size_t)s->team_nblk * gr.nown);
        int rr = pb200_solver_get_state(s->parts[r], loc.data());
        if (!rr) team_paste(gr, nt, loc, s->team_nblk, out);
        return rr;
    });
}
static int team_solver_step(pb200_solver *s, const pb200_step_in *in, const pb200_krylov_opts *opts, pb200_step_stats *stats)
{
    pb200_ctx *tc = s->ctx;
    const int n = tc->team->n;
    std::vector<pb200_step_stats> st(n);
    std::vector<int> rcs(n, 0);
    int rc = team_run(tc, [&](int r) {
        const Grid &gr = s->parts[r]->g;
        const size_t off = (size_t)gr.k0 * gr.plane;
        pb200_step_in ir = *in;
        for (int p = 0; p < 2; ++p) for (int w = 0; w < 2; ++w) if (in->f_arr[p][w]) ir.f_arr[p][w] = in->f_arr[p][w] + off;
        for (int w = 0; w < 2; ++w) if (in->g_arr[w]) ir.g_arr[w] = in->g_arr[w] + off;
        memset(&st[r], 0, sizeof(st[r]));
        rcs[r] = pb200_solver_step(s->parts[r], &ir, opts, &st[r]);
        return rcs[r] == PB200_ENOTCONV ? PB200_OK : rcs[r];     // (every rank takes the same decision; reported below)
    });
    if (rc) return rc;
    if (stats) {
        *stats = st[0];      // iteration counts, norms, global DOF counts are identical on every rank
        for (int r = 1; r < n; ++r) {
            if (st[r].solve_ms > stats->solve_ms) stats->solve_ms = st[r].solve_ms;
            if (st[r].setup_ms > stats->setup_ms) stats->setup_ms = st[r].setup_ms;
            stats->launches += st[r].launches; stats->apply_launches += st[r].apply_launches;
            stats->apply_cells_uniform += st[r].apply_cells_uniform; stats->apply_cells_general += st[r].apply_cells_general;
            stats->apply_cells_fast += st[r].apply_cells_fast; stats->band_cells += st[r].band_cells; stats->band_rows += st[r].band_rows;
        }
    }
    if (rcs[0] == PB200_ENOTCONV) return set_err(tc, PB200_ENOTCONV, "Krylov solve did not reach the tolerance within maxit iterations");
    return PB200_OK;
}
static int team_solver_norms(pb200_solver *s, int phase, const double *u_ana, double p, int relative, double *out)
{
    pb200_ctx *tc = s->ctx;
    std::vector<double> o((size_t)4 * tc->team->n);
    int rc = team_run(tc, [&](int r) {
        const Grid &gr = s->parts[r]->g;
        return pb200_solver_error_norms(s->parts[r], phase, u_ana + (size_t)gr.k0 * gr.plane, p, relative, o.data() + 4 * r);
    });
    if (!rc) memcpy(out, o.data(), 4 * sizeof(double));      // (reduced over the ranks inside the library: every rank holds the same numbers)
    return rc;
}
