// assemble.cuh -- per-step right-hand sides, known-value elimination and state write-back.
// Restates b_mono_unstead_diff / b_diph_unstead_diff / b_*_stead_diff (/root/reference/src/solver/diffusion.jl:45-58,
// 146-161, 243-265, 391-420), the border rows of BC_border_mono!/diph! (src/solver.jl:417-580) and the scatter of
// solve_system! (src/solver.jl:186-187) on the device, fused into one pass per step.
#pragma once
#include "operators.cuh"

struct SrcSpec { const double *arr; double cst; };   // value(l) = arr ? arr[l] : cst
__device__ __forceinline__ double src_at(const SrcSpec &s, int64_t l) { return s.arr ? s.arr[l] : s.cst; }

struct StepCoef {
    double cV;     // 1 unsteady / 0 steady
    double c;      // implicit coefficient on D G'W!(..): dt (BE), dt/2 (CN), 1 (steady)
    double ce;     // explicit coefficient: dt/2 (CN) else 0
    double c2;     // mono interface row scale: dt/2 (CN) else 1
    double wf0, wf1;  // source weights: V * (wf0 f(t_n) + wf1 f(t_n+dt))  -- BE (0,dt) CN (dt/2,dt/2) steady (1,0)
    double wg0, wg1;  // mono interface data weights: g_eff = wg0 g0 + wg1 g1 -- BE (0,1) CN (1,1) steady (1,0)
    int cn;        // 1: Crank-Nicolson
    int sym;       // symmetrised rows (CG path)
};

// mono, Dirichlet interface: kept T_gamma values. BE/steady: g ; CN: g0 + g1 - T_gamma^n (row 2 of diffusion.jl:227,258)
__global__ void k_gamma_known(Grid g, StepCoef sc, const unsigned char *__restrict__ m, SrcSpec g0, SrcSpec g1, const double *__restrict__ Tg,
                              double *__restrict__ gK)
{
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (m[l] & MB_IKNOWN) {
            v = sc.wg0 * src_at(g0, l) + sc.wg1 * src_at(g1, l);
            if (sc.cn) v -= Tg[l];
        }
        gK[l] = v;
    }
}

template <int N>
__global__ void k_rhs_mono(Grid g, PhaseDev p, SysParams sp, StepCoef sc, const unsigned char *__restrict__ m, const double *__restrict__ Tw,
                           const double *__restrict__ Tg, const double *__restrict__ ufix, const double *__restrict__ gK, SrcSpec f0, SrcSpec f1,
                           SrcSpec g0, SrcSpec g1, double *__restrict__ bb, double *__restrict__ bi, int skip_known,
                           const long long *__restrict__ list = nullptr, int nlist = 0)
{
    // list != nullptr: only the listed rows (the folded path's CN step evaluates the unfolded explicit part on the rows that couple to
    // eliminated values and takes M^ x^n everywhere else, kf_rhs_dense)
    const int64_t nwork = list ? (int64_t)nlist : g.nown;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwork; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = list ? (int64_t)list[w] : w + g.plane, t = l - g.plane;
        const unsigned char mb = m[l];
        const bool wb = mb & MB_FREE, wi = bi && (mb & MB_IFREE);
        if (!wb && !wi) { bb[l] = 0.0; if (bi) bi[l] = 0.0; continue; }
        int c[PB_MAXD] = {0, 0, 0};
        if (sc.cn || !skip_known) cell_coords(g, t, c);
        const double D = D_at(p, l), V = p.V[l], Gm = p.Gam[l];
        double Rbk = 0.0, Rik = 0.0, Rbe = 0.0, Rie = 0.0;
        GamSpec gk = {gK, 1.0, nullptr, 0.0, 0.0};
        // zero for every other row: nothing known is within reach; skip_known: k_rhs_known_mono adds it for the listed rows afterwards
        double Cvk = 0.0, Cve = 0.0;   // advective parts (conv_row, operators.cuh): known values / explicit CN part
        if (!skip_known && (wi || (mb & MB_KNBR))) { phase_rows<N>(p, g, l, c, ufix, gk, Rbk, Rik); Cvk = conv_row<N>(p, g, l, c, ufix, gk); }
        if (sc.cn) {
            GamSpec ge = {Tg, 1.0, nullptr, 0.0, 0.0};
            phase_rows<N>(p, g, l, c, Tw, ge, Rbe, Rie);
            Cve = conv_row<N>(p, g, l, c, Tw, ge);
        }
        if (wb) {
            double v = sc.cV * V * Tw[l] + V * (sc.wf0 * src_at(f0, l) + sc.wf1 * src_at(f1, l)) - sc.ce * (D * Rbe + Cve)
                       - sc.c * (D * Rbk + Cvk);   // (ufix is zero on a free row)
            if (sc.sym) v /= D;
            bb[l] = v;
        } else bb[l] = 0.0;
        if (bi) {
            if (wi) {
                const double gKl = gK ? gK[l] : 0.0;
                double v = sc.c2 * Gm * (sc.wg0 * src_at(g0, l) + sc.wg1 * src_at(g1, l))
                           - sc.ce * (sp.beta * Rie + sp.alpha * Gm * (sc.cn ? Tg[l] : 0.0))
                           - sc.c2 * (sp.beta * Rik + sp.alpha * Gm * gKl);
                if (sc.sym) v *= sc.c / (sc.c2 * sp.beta);
                bi[l] = v;
            } else bi[l] = 0.0;
        }
    }
}

// diph: known gamma1 = g / a1 everywhere (row 2), gamma2 known part = 0
template <int N>
__global__ void k_rhs_diph(Grid g, PhaseDev p1, PhaseDev p2, SysParams sp, StepCoef sc, const unsigned char *__restrict__ m1,
                           const unsigned char *__restrict__ m2, const double *__restrict__ Tw1, const double *__restrict__ Tg1,
                           const double *__restrict__ Tw2, const double *__restrict__ Tg2, const double *__restrict__ ufix1,
                           const double *__restrict__ ufix2, SrcSpec f10, SrcSpec f11, SrcSpec f20, SrcSpec f21, SrcSpec gj, SrcSpec hj,
                           double *__restrict__ b1, double *__restrict__ b2, double *__restrict__ bw, int skip_known,
                           const long long *__restrict__ list = nullptr, int nlist = 0)
{
    const int64_t nwork = list ? (int64_t)nlist : g.nown;   // (list: see k_rhs_mono)
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwork; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = list ? (int64_t)list[w] : w + g.plane, t = l - g.plane;
        const unsigned char a = m1[l], b = m2[l];
        const bool w1 = a & MB_FREE, w2 = b & MB_FREE, ww = b & MB_IFREE;
        if (!w1 && !w2 && !ww) { b1[l] = 0.0; b2[l] = 0.0; bw[l] = 0.0; continue; }
        int c[PB_MAXD] = {0, 0, 0};
        if (sc.cn || !skip_known) cell_coords(g, t, c);   // (64-bit divisions: only when a stencil is evaluated; the BE fast path streams)
        GamSpec gk1 = {gj.arr, 1.0 / sp.a1, nullptr, 0.0, gj.arr ? 0.0 : gj.cst / sp.a1};
        GamSpec gk2 = {nullptr, 0.0, nullptr, 0.0, 0.0};
        double Rbk1 = 0, Rik1 = 0, Rbk2 = 0, Rik2 = 0, Rbe1 = 0, Rie1 = 0, Rbe2 = 0, Rie2 = 0;
        // the known part is zero unless something known is within the row's reach (MB_KNBR) -- or the row is an interface row
        double Ck1 = 0, Ck2 = 0, Ce1 = 0, Ce2 = 0;   // advective parts (conv_row, operators.cuh)
        if (!skip_known) {
            if ((w1 && (a & MB_KNBR)) || ww) { phase_rows<N>(p1, g, l, c, ufix1, gk1, Rbk1, Rik1); if (w1) Ck1 = conv_row<N>(p1, g, l, c, ufix1, gk1); }
            if ((w2 && (b & MB_KNBR)) || ww) { phase_rows<N>(p2, g, l, c, ufix2, gk2, Rbk2, Rik2); if (w2) Ck2 = conv_row<N>(p2, g, l, c, ufix2, gk2); }
        }
        if (sc.cn) {
            GamSpec ge1 = {Tg1, 1.0, nullptr, 0.0, 0.0}, ge2 = {Tg2, 1.0, nullptr, 0.0, 0.0};
            if (w1) { phase_rows<N>(p1, g, l, c, Tw1, ge1, Rbe1, Rie1); Ce1 = conv_row<N>(p1, g, l, c, Tw1, ge1); }
            if (w2) { phase_rows<N>(p2, g, l, c, Tw2, ge2, Rbe2, Rie2); Ce2 = conv_row<N>(p2, g, l, c, Tw2, ge2); }
        }
        const double V1 = p1.V[l], V2 = p2.V[l];
        b1[l] = w1 ? sc.cV * V1 * Tw1[l] + V1 * (sc.wf0 * src_at(f10, l) + sc.wf1 * src_at(f11, l)) - sc.ce * (D_at(p1, l) * Rbe1 + Ce1)
                         - sc.c * (D_at(p1, l) * Rbk1 + Ck1)       /* (ufix is zero on a free row) */
                   : 0.0;
        b2[l] = w2 ? sc.cV * V2 * Tw2[l] + V2 * (sc.wf0 * src_at(f20, l) + sc.wf1 * src_at(f21, l)) - sc.ce * (D_at(p2, l) * Rbe2 + Ce2)
                         - sc.c * (D_at(p2, l) * Rbk2 + Ck2)
                   : 0.0;
        bw[l] = ww ? p2.Gam[l] * src_at(hj, l) - (sp.b1 * Rik1 + sp.b2 * Rik2) : 0.0;
    }
}

// initial guess: previous state restricted to the free sets (warm) or zero
// The "known part" of the right-hand side (couplings of a row to eliminated values) for the compact, sorted list of rows that have
// one (MB_KNBR rows and interface rows): subtracts it from the b written by k_rhs_* with skip_known = 1.  One thread per listed cell --
// the heavy unfolded stencil evaluations run side by side instead of stalling the warps of the streaming kernel.
template <int N>
__global__ void k_rhs_known_mono(Grid g, PhaseDev p, SysParams sp, StepCoef sc, const unsigned char *__restrict__ m, const long long *__restrict__ list, int n,
                                 const double *__restrict__ ufix, const double *__restrict__ gK, double *__restrict__ bb, double *__restrict__ bi,
                                 int set_ifc = 0, SrcSpec g0 = SrcSpec{nullptr, 0.0}, SrcSpec g1 = SrcSpec{nullptr, 0.0})
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t l = list[i];
        int c[PB_MAXD];
        cell_coords(g, l - g.plane, c);
        const unsigned char mb = m[l];
        const bool wb = mb & MB_FREE, wi = bi && (mb & MB_IFREE);
        if (!wb && !wi) continue;
        double Rbk, Rik;
        GamSpec gk = {gK, 1.0, nullptr, 0.0, 0.0};
        phase_rows<N>(p, g, l, c, ufix, gk, Rbk, Rik);
        if (wb) { double v = sc.c * (D_at(p, l) * Rbk + conv_row<N>(p, g, l, c, ufix, gk)); if (sc.sym) v /= D_at(p, l); bb[l] -= v; }
        if (wi) {
            double v = sc.c2 * sp.beta * Rik;
            if (sc.sym) v *= sc.c / (sc.c2 * sp.beta);
            // set_ifc (steps without an explicit part, kf_rhs_dense wrote the bulk rows only): the interface row from scratch
            const double base = set_ifc ? sc.c2 * p.Gam[l] * (sc.wg0 * src_at(g0, l) + sc.wg1 * src_at(g1, l)) - sc.c2 * sp.alpha * p.Gam[l] * (gK ? gK[l] : 0.0)
                                        : bi[l];
            bi[l] = base - v;
        }
    }
}
template <int N>
__global__ void k_rhs_known_diph(Grid g, PhaseDev p1, PhaseDev p2, SysParams sp, StepCoef sc, const unsigned char *__restrict__ m1, const unsigned char *__restrict__ m2,
                                 const long long *__restrict__ list, int n, const double *__restrict__ ufix1, const double *__restrict__ ufix2, SrcSpec gj,
                                 double *__restrict__ b1, double *__restrict__ b2, double *__restrict__ bw, int set_ifc = 0,
                                 SrcSpec hj = SrcSpec{nullptr, 0.0})
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t l = list[i];
        int c[PB_MAXD];
        cell_coords(g, l - g.plane, c);
        const unsigned char a = m1[l], b = m2[l];
        const bool w1 = a & MB_FREE, w2 = b & MB_FREE, ww = b & MB_IFREE;
        GamSpec gk1 = {gj.arr, 1.0 / sp.a1, nullptr, 0.0, gj.arr ? 0.0 : gj.cst / sp.a1};
        GamSpec gk2 = {nullptr, 0.0, nullptr, 0.0, 0.0};
        double Rbk1 = 0, Rik1 = 0, Rbk2 = 0, Rik2 = 0;
        if ((w1 && (a & MB_KNBR)) || ww) phase_rows<N>(p1, g, l, c, ufix1, gk1, Rbk1, Rik1);
        if ((w2 && (b & MB_KNBR)) || ww) phase_rows<N>(p2, g, l, c, ufix2, gk2, Rbk2, Rik2);
        if (w1) b1[l] -= sc.c * (D_at(p1, l) * Rbk1 + ((a & MB_KNBR) ? conv_row<N>(p1, g, l, c, ufix1, gk1) : 0.0));
        if (w2) b2[l] -= sc.c * (D_at(p2, l) * Rbk2 + ((b & MB_KNBR) ? conv_row<N>(p2, g, l, c, ufix2, gk2) : 0.0));
        if (ww) bw[l] = (set_ifc ? p2.Gam[l] * src_at(hj, l) : bw[l]) - (sp.b1 * Rik1 + sp.b2 * Rik2);
    }
}
// rows that need the known part: MB_KNBR bulk rows and every interface row
__global__ void k_mark_known_rows(Grid g, const unsigned char *__restrict__ m1, const unsigned char *__restrict__ m2, long long *list, int *count, int cap)
{
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char a = m1[l], b = m2 ? m2[l] : 0;
        if ((a | b) & (MB_KNBR | MB_IFREE)) { const int k = atomicAdd(count, 1); if (k < cap) list[k] = l; }
    }
}

// Initial guess by polynomial extrapolation in time: x = sum_j c_j T^(n-j), m = number of states used (0: zero guess, 1: T^n,
// 2: 2 T^n - T^(n-1), 3: 3, -3, 1, ...: the unique polynomial of degree m-1 through the last m states, evaluated one step ahead).
#define PB_MAXHIST 5
struct GuessSpec { int m; const double *T[PB_MAXHIST]; double c[PB_MAXHIST]; };
__global__ void k_guess(Grid g, GuessSpec gs, const unsigned char *__restrict__ m, unsigned char bit, double *__restrict__ x)
{
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (gs.m && (m[l] & bit)) {
#pragma unroll
            for (int j = 0; j < PB_MAXHIST; ++j)
                if (j < gs.m) v += gs.c[j] * gs.T[j][l];
        }
        x[l] = v;
    }
}

// state write-back (solve_system!: removed DOFs are exactly 0; border rows hold their value)
__global__ void k_store_bulk(Grid g, const double *__restrict__ x, const double *__restrict__ ufix, double *__restrict__ T)
{
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x)
        T[l] = x[l] + ufix[l];
}
// diph: T_gamma2 = w ; T_gamma1 = (g + a2 w) / a1 on every cell
__global__ void k_store_diph_ifc(Grid g, SysParams sp, SrcSpec gj, const double *__restrict__ w, double *__restrict__ Tg1, double *__restrict__ Tg2)
{
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        const double wv = w[l];
        Tg2[l] = wv;
        Tg1[l] = (src_at(gj, l) + sp.a2 * wv) / sp.a1;
    }
}

// DOF census: [0] bulk kept (free + fixed, all phases), [1] interface kept
__global__ void k_count(Grid g, const unsigned char *__restrict__ m1, const unsigned char *__restrict__ m2, double *partials, double *results,
                        unsigned *counter)
{
    double v[2] = {0.0, 0.0};
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char a = m1[l], b = m2 ? m2[l] : 0;
        v[0] += ((a & (MB_FREE | MB_FIXED | MB_SLAVE)) ? 1.0 : 0.0) + ((b & (MB_FREE | MB_FIXED | MB_SLAVE)) ? 1.0 : 0.0);
        v[1] += ((a & (MB_IFREE | MB_IKNOWN)) ? 1.0 : 0.0) + ((b & MB_IFREE) ? 1.0 : 0.0);
    }
    block_reduce_publish<2>(v, partials, results, counter);
}


// check_convergence (src/convergence.jl:4-93): volume-weighted error norms of u_ana - T_omega by cell class (full / cut / empty), ONE fused
// reduction on the device state.  Classes by capacity.cell_types (1, -1, 0).  sums[c] = sum_class |e|^p V (finite p; e = err or err / u_ana),
// or sum_class u_ana^2 (p = Inf, relative); sums[3] = sum(V) over every cell; mx[c] = max_class |err|, mx[3 + c] = max_class |u_ana| (bit
// patterns of non-negative doubles order like unsigned integers, so atomicMax keeps the result independent of the block order).
__global__ void k_err_norms(Grid g, const double *__restrict__ ct, const double *__restrict__ V, const double *__restrict__ ua,
                            const double *__restrict__ Tw, double p, int relative, int is_inf, unsigned long long *__restrict__ mx,
                            double *partials, double *results, unsigned *counter)
{
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    double me[3] = {0.0, 0.0, 0.0}, mu[3] = {0.0, 0.0, 0.0};
    for (int64_t l = g.plane + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (int64_t)gridDim.x * blockDim.x) {
        const double c = ct[l], Vi = V[l], u = ua[l], e = u - Tw[l];
        v[3] += Vi;
        const int cls = c == 1.0 ? 0 : (c == -1.0 ? 1 : (c == 0.0 ? 2 : -1));
        if (cls < 0) continue;
        double add = 0.0;
        if (!is_inf) add = pow(fabs(relative ? e / u : e), p) * Vi;
        else if (relative) add = u * u;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k == cls) { v[k] += add; me[k] = fmax(me[k], fabs(e)); mu[k] = fmax(mu[k], fabs(u)); }
    }
    if (is_inf) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { me[k] = fmax(me[k], __shfl_xor_sync(0xffffffffu, me[k], o)); mu[k] = fmax(mu[k], __shfl_xor_sync(0xffffffffu, mu[k], o)); }
            if ((threadIdx.x & 31) == 0) {
                atomicMax(mx + k, (unsigned long long)__double_as_longlong(me[k]));
                atomicMax(mx + 3 + k, (unsigned long long)__double_as_longlong(mu[k]));
            }
        }
    }
    block_reduce_publish<4>(v, partials, results, counter);
}
