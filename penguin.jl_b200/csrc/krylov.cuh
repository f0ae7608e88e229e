// krylov.cuh -- Jacobi-preconditioned CG (symmetric monophasic system) and BiCGSTAB (diphasic / general system).
// Replaces the linear solve of /root/reference/src/solver.jl:158-188 (UMFPACK `\` or IterativeSolvers gmres/cg/bicgstabl)
// on the reduced system that remove_zero_rows_cols! (src/solver.jl:59-78) would produce.
//
// All scalars (rho, alpha, omega, ...) stay on the device: reductions publish into ctx->d_results and the consumer
// kernels recompute alpha = rho / sigma from those slots, so the host only synchronises when it tests convergence.
// Krylov vectors are multi-field (bulk field(s) + interface field); they are exactly zero outside the free sets, an
// invariant kept by the masked operator and the masked right-hand side, so the vector kernels need no masks.
#pragma once
#include "common.cuh"

#define KV_MAXF 3
struct MVec { double *f[KV_MAXF]; };

// result slots
// (rho, rr) live in two ping-pong pairs {0,1} and {2,3}: a reduction publishes the NEW pair while consumers still read the old rho
enum { SL_PAIR0 = 0, SL_PAIR1 = 2, SL_SIGMA = 4, SL_TS = 5, SL_TT = 6, SL_BB = 7, SL_TMP = 8 };

__device__ __forceinline__ double safe_div(double a, double b) { return b != 0.0 ? a / b : 0.0; }

#define KV_LOOP(g)                                                                                                   \
    const int fld = blockIdx.y;                                                                                      \
    for (int64_t l = (g).plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < (g).plane + (g).nown;           \
         l += (int64_t)gridDim.x * blockDim.x)

// generic: up to two dots  res[s0] = (a,b), res[s1] = (c,d) over all fields  (launched with gridDim.y == 1, loops fields)
template <int K>
__global__ void k_dots(Grid g, int nf, MVec a, MVec b, MVec c, MVec d, double *partials, double *results, unsigned *counter)
{
    double v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = 0.0;
    for (int f = 0; f < nf; ++f) {
        const double *__restrict__ pa = a.f[f], *__restrict__ pb = b.f[f];
        const double *__restrict__ pc = K > 1 ? c.f[f] : nullptr, *__restrict__ pd = K > 1 ? d.f[f] : nullptr;
        for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
            v[0] += pa[l] * pb[l];
            if (K > 1) v[1] += pc[l] * pd[l];
        }
    }
    block_reduce_publish<K>(v, partials, results, counter);
}

// y = a*x + b*y (per field; scalars immediate)
__global__ void k_axpby(Grid g, double a, MVec x, double b, MVec y)
{
    KV_LOOP(g) { y.f[fld][l] = a * x.f[fld][l] + b * y.f[fld][l]; }
}
__global__ void k_copy(Grid g, MVec x, MVec y)
{
    KV_LOOP(g) { y.f[fld][l] = x.f[fld][l]; }
}
// y = dinv .* x
__global__ void k_scale(Grid g, MVec dinv, MVec x, MVec y)
{
    KV_LOOP(g) { y.f[fld][l] = dinv.f[fld][l] * x.f[fld][l]; }
}
// r = b - q ; publishes (r,r) [K=1] -- initial residual
__global__ void k_resid(Grid g, int nf, MVec b, MVec q, MVec r, double *partials, double *results, unsigned *counter)
{
    double v[1] = {0.0};
    for (int f = 0; f < nf; ++f)
        for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
            const double x = b.f[f][l] - q.f[f][l];
            r.f[f][l] = x;
            v[0] += x * x;
        }
    block_reduce_publish<1>(v, partials, results, counter);
}

// ---- CG ------------------------------------------------------------------------------------------------------
// x += alpha p ; r -= alpha q ; publishes rho_new = (r, dinv r), rr = (r, r)   with alpha = rho / sigma
__global__ void k_cg_update(Grid g, int nf, double *res, int sl_rho, int sl_rho_new, MVec dinv, MVec p, MVec q, MVec x, MVec r,
                            double *partials, unsigned *counter)
{
    const double alpha = safe_div(res[sl_rho], res[SL_SIGMA]);
    double v[2] = {0.0, 0.0};
    for (int f = 0; f < nf; ++f)
        for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
            x.f[f][l] += alpha * p.f[f][l];
            const double rn = r.f[f][l] - alpha * q.f[f][l];
            r.f[f][l] = rn;
            v[0] += rn * dinv.f[f][l] * rn;
            v[1] += rn * rn;
        }
    block_reduce_publish<2>(v, partials, res + sl_rho_new, counter);   // (rho_new, rr) pair
}
// p = dinv r + beta p, beta = rho_new / rho
__global__ void k_cg_p(Grid g, const double *res, int sl_rho, int sl_rho_new, MVec dinv, MVec r, MVec p)
{
    const double beta = safe_div(res[sl_rho_new], res[sl_rho]);
    KV_LOOP(g) { p.f[fld][l] = dinv.f[fld][l] * r.f[fld][l] + beta * p.f[fld][l]; }
}
// ---- BiCGSTAB --------------------------------------------------------------------------------------------------
// s = r - alpha v ; sh = dinv s          (alpha = rho / sigma)
__global__ void k_bicg_s(Grid g, const double *res, int sl_rho, MVec dinv, MVec r, MVec v, MVec s, MVec sh)
{
    const double alpha = safe_div(res[sl_rho], res[SL_SIGMA]);
    KV_LOOP(g) {
        const double x = r.f[fld][l] - alpha * v.f[fld][l];
        s.f[fld][l] = x;
        sh.f[fld][l] = dinv.f[fld][l] * x;
    }
}
// x += alpha ph + omega sh ; r = s - omega t ; publishes the new pair rho_new = (r0, r), rr = (r, r)
__global__ void k_bicg_xr(Grid g, int nf, double *res, int sl_rho, int sl_rho_new, MVec ph, MVec sh, MVec s, MVec t, MVec r0, MVec x, MVec r, double *partials,
                          unsigned *counter)
{
    const double alpha = safe_div(res[sl_rho], res[SL_SIGMA]);
    const double omega = safe_div(res[SL_TS], res[SL_TT]);
    double v[2] = {0.0, 0.0};
    for (int f = 0; f < nf; ++f)
        for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
            x.f[f][l] += alpha * ph.f[f][l] + omega * sh.f[f][l];
            const double rn = s.f[f][l] - omega * t.f[f][l];
            r.f[f][l] = rn;
            v[0] += r0.f[f][l] * rn;
            v[1] += rn * rn;
        }
    block_reduce_publish<2>(v, partials, res + sl_rho_new, counter);   // (rho_new, rr) pair
}
// p = r + beta (p - omega v) ; ph = dinv p     beta = (rho_new / rho) (alpha / omega)
__global__ void k_bicg_p(Grid g, const double *res, int sl_rho, int sl_rho_new, MVec dinv, MVec r, MVec v, MVec p, MVec ph)
{
    const double alpha = safe_div(res[sl_rho], res[SL_SIGMA]);
    const double omega = safe_div(res[SL_TS], res[SL_TT]);
    const double beta = safe_div(res[sl_rho_new], res[sl_rho]) * safe_div(alpha, omega);
    KV_LOOP(g) {
        const double x = r.f[fld][l] + beta * (p.f[fld][l] - omega * v.f[fld][l]);
        p.f[fld][l] = x;
        ph.f[fld][l] = dinv.f[fld][l] * x;
    }
}
// dinv = 1 / diag (entries of non-free rows were set to 1 by the diagonal kernels)
__global__ void k_recip(Grid g, MVec d)
{
    KV_LOOP(g) { d.f[fld][l] = 1.0 / d.f[fld][l]; }
}
