// operators.cuh -- matrix-free G / H / W-dagger stencils (the sparse products of /root/reference/src/operators.jl:127-158
// and the block rows of /root/reference/src/solver/diffusion.jl:30-43, 104-144, 212-241, 334-389 are never assembled).
//
// Per phase and direction d, with i-1 the neighbour in direction d (checked against the sparse
// definition by tests/test_operators.py):
//     e_i      = 0 on the last padded index of direction d, else 1          (delta_m zeroes D[n,n], operators.jl:9)
//     q_{d,i}  = W!_{d,i} [ e_i B_i u_i - B_{i-1} u_{i-1} + e_i (A_i - B_i) g_i - (A_i - B_{i-1}) g_{i-1} ]
//     (G'q)_i  = B_i ( e_i q_{d,i} - q_{d,i+1} )
//     (H'q)_i  = e_i (A_i - B_i) q_{d,i} - (A_{i+1} - B_i) q_{d,i+1}
// Faces i = 0 (no i-1 terms) and i = n_d (W! = 1 where W = 0) are included exactly as the reference includes them.
#pragma once
#include "common.cuh"

struct PhaseDev {
    const double *V, *Gam;
    const double *A[PB_MAXD], *B[PB_MAXD], *Wd[PB_MAXD];
    const double *Darr;  // per-cell D or nullptr
    double Dc;           // constant D when Darr == nullptr
    // advection (ConvectionOps, /root/reference/src/operators.jl:194-209): nullptr without it
    const double *cf[PB_MAXD];   // cf_d = S_m (A_d u_d): the face flux coefficient of C_d = D_p diag(cf_d) S_m
    const double *kd;            // 0.5 sum_d S_p^(d) (H' u_gamma): the diagonal of 0.5 sum_d K_d
};

// gamma "spec": value(l) = scale * (ptr ? ptr[l] : 0) + (off_ptr ? off_scale * off_ptr[l] : 0) + cst
struct GamSpec {
    const double *ptr; double scale;
    const double *off_ptr; double off_scale;
    double cst;
};
__device__ __forceinline__ double gam_at(const GamSpec &s, int64_t l)
{
    double v = s.cst;
    if (s.ptr) v += s.scale * s.ptr[l];
    if (s.off_ptr) v += s.off_scale * s.off_ptr[l];
    return v;
}

// raw rows of one phase at local index l (global coords c): Rb = [G' W! (G u + H g)]_l,  Ri = [H' W! (G u + H g)]_l
template <int N>
__device__ __forceinline__ void phase_rows(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD], const double *__restrict__ u,
                                           const GamSpec &gs, double &Rb, double &Ri)
{
    Rb = 0.0; Ri = 0.0;
    const double ui = u ? u[l] : 0.0;
    const double gi = gam_at(gs, l);
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int64_t s = g.stride[d];
        const int id = c[d], last = g.pd[d] - 1;
        const double *__restrict__ A = ph.A[d], *__restrict__ B = ph.B[d], *__restrict__ W = ph.Wd[d];
        const double bi = B[l], ai = A[l];
        const double ei = id < last ? 1.0 : 0.0;
        // lower face (d, i)
        double qL = ei * (bi * ui + (ai - bi) * gi);
        if (id > 0) {
            const double bm = B[l - s];
            qL -= bm * (u ? u[l - s] : 0.0) + (ai - bm) * gam_at(gs, l - s);
        }
        qL *= W[l];
        // upper face (d, i+1)
        double qU = 0.0, hpU = 0.0;
        if (id < last) {
            const double bp = B[l + s], ap = A[l + s];
            const double ep = id + 1 < last ? 1.0 : 0.0;
            hpU = ap - bi;
            qU = W[l + s] * (ep * (bp * (u ? u[l + s] : 0.0) + (ap - bp) * gam_at(gs, l + s)) - bi * ui - hpU * gi);
        }
        Rb += bi * (ei * qL - qU);
        Ri += ei * (ai - bi) * qL - hpU * qU;
    }
}

// advective part of a bulk row (A_mono_unstead_advdiff, /root/reference/src/solver/advectiondiffusion.jl:178-210):
//   [(sum_d C_d + 0.5 sum_d K_d) u + 0.5 sum_d K_d gamma]_l,   (C_d u)_l = w_{l+s} - w_l  (0 on the last padded index),
//   w_j = cf_j (S_m u)_j,  (S_m u)_j = (u_j + u_{j-1}) / 2  below the last index,  u_{j-1} / 2  on it  (delta_p, Sigma_m of src/operators.jl:8-12)
template <int N>
__device__ __forceinline__ double conv_row(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD], const double *__restrict__ u, const GamSpec &gs)
{
    if (!ph.kd) return 0.0;
    const double ui = u ? u[l] : 0.0;
    double r = ph.kd[l] * (ui + gam_at(gs, l));
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int64_t s = g.stride[d];
        const int id = c[d], last = g.pd[d] - 1;
        if (id >= last) continue;                               // delta_p: the last row is zero
        const double um = (id > 0 && u) ? u[l - s] : 0.0, up = u ? u[l + s] : 0.0;
        const double wl = ph.cf[d][l] * 0.5 * (ui + um);        // (id < last: the regular Sigma_m row)
        const double wu = ph.cf[d][l + s] * 0.5 * (id + 1 < last ? up + ui : ui);
        r += wu - wl;
    }
    return r;
}
template <int N>
__device__ __forceinline__ double conv_diag(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD])
{
    if (!ph.kd) return 0.0;
    double r = ph.kd[l];
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int id = c[d], last = g.pd[d] - 1;
        if (id < last) r += 0.5 * (ph.cf[d][l + g.stride[d]] - ph.cf[d][l]);
    }
    return r;
}
// remove_zero_rows_cols! (src/solver.jl:59-78) keeps an unknown whose row AND column of the assembled matrix are non-zero.  The advective operators are
// not symmetric and reach one cell further than the capacities of a cell say: cf_j = (a_j + a_{j-1}) / 2 is non-zero on a SOLID cell j whose lower
// neighbour is fluid, so the reference keeps T_j of that solid cell as an unknown (row: -w_j = 0, i.e. T_j = -T_{j-1}).  This predicate reproduces it.
template <int N>
__device__ __forceinline__ bool conv_keeps(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD])
{
    if (!ph.kd) return false;
    double rs = 0.0, cs = 0.0, dg = ph.kd[l];
    rs += fabs(ph.kd[l]);                       // (the T_gamma column of the row; counted although that column may be trimmed: kd != 0 only beside the interface)
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int64_t s = g.stride[d];
        const int id = c[d], last = g.pd[d] - 1;
        if (id < last) {
            const double cl = ph.cf[d][l], cu = ph.cf[d][l + s];
            dg += 0.5 * (cu - cl);
            if (id + 1 < last) rs += fabs(0.5 * cu);      // coefficient of u[l + s] in row l
            if (id > 0) rs += fabs(0.5 * cl);             // coefficient of u[l - s] in row l
            // column l: rows l + s (exists below the last index) and l - s
            if (id + 1 < last) cs += fabs(0.5 * cu);      // u[l] in w_{l+s}, row l + s:  -w_{l+s}
            if (id > 0) cs += fabs(0.5 * cl);             // u[l] in w_l, row l - s:  +w_l
        }
    }
    rs += fabs(dg); cs += fabs(dg);
    return rs != 0.0 && cs != 0.0;
}
// ConvectionOps set-up: cf_d = S_m (A_d u_d), kd = 0.5 sum_d S_p^(d) q with q = H' u_gamma (computed by k_div)
template <int N>
__global__ void k_conv_coef(Grid g, PhaseDev ph, const double *__restrict__ uo, const double *__restrict__ q, double *__restrict__ cf, double *__restrict__ kd)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        double k = 0.0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            const int64_t s = g.stride[d];
            const int id = c[d], last = g.pd[d] - 1;
            const double *__restrict__ A = ph.A[d];
            const double *__restrict__ ud = uo + (int64_t)d * g.nloc;
            const double am = id > 0 ? A[l - s] * ud[l - s] : 0.0;
            cf[(int64_t)d * g.nloc + l] = id < last ? 0.5 * (A[l] * ud[l] + am) : 0.5 * am;      // Sigma_m: [n, n] = 0
            if (id < last) k += 0.5 * (q[l] + q[l + s]);                                           // Sigma_p: last row zero
        }
        kd[l] = 0.5 * k;
    }
}

// diagonal entries: GG_ll = sum_d B_l^2 (e_l W!_l + W!_{l+1}),  HH_ll = sum_d e_l (A_l-B_l)^2 W!_l + (A_{l+1}-B_l)^2 W!_{l+1}
// also the "row of H' is nonzero" predicate used by the reference's zero-row trimming (src/solver.jl:59-78)
template <int N>
__device__ __forceinline__ void phase_diag(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD], double &GG, double &HH, bool &rowG,
                                           bool &hrow)
{
    GG = 0.0; HH = 0.0; rowG = false; hrow = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int64_t s = g.stride[d];
        const int id = c[d], last = g.pd[d] - 1;
        const double bi = ph.B[d][l], ai = ph.A[d][l];
        const double ei = id < last ? 1.0 : 0.0;
        const double hm = ei * (ai - bi);
        double wU = 0.0, hp = 0.0;
        if (id < last) { wU = ph.Wd[d][l + s]; hp = ph.A[d][l + s] - bi; }
        const double wL = ph.Wd[d][l];
        GG += bi * bi * (ei * wL + wU);
        HH += hm * hm * wL + hp * hp * wU;
        rowG = rowG || (bi != 0.0);
        hrow = hrow || (hm != 0.0) || (hp != 0.0);
    }
}

// ------------------------------------------------------------------------------------------------------------
// mask bits (one byte per local cell and phase)
// ------------------------------------------------------------------------------------------------------------
#define MB_FREE 1   // bulk unknown solved for
#define MB_FIXED 2  // bulk unknown pinned by a border Dirichlet row
#define MB_IFREE 4  // interface unknown solved for (mono Robin/Neumann: T_gamma ; diph: T_gamma2)
#define MB_IKNOWN 8 // mono Dirichlet interface: T_gamma = g kept by the reference's trimming (Gamma != 0)
#define MB_SLAVE 32 // 1-D Neumann border row (src/solver.jl:471-493): (x_row - x_adj) / dx = g, i.e. x_row = x_adj + g dx is eliminated: the value
                    // follows its neighbour (k_slave_copy around every operator apply), g dx sits in ufix
#define MB_KNBR 16  // the bulk row of this cell couples to an eliminated value (a known T_gamma through H, or a border-Dirichlet neighbour):
                    // only these rows need the "known part" of the right-hand side (assemble.cuh)

struct BorderDev {
    int kind[6];
    int present[6];           // the key exists in BorderConditions (whatever its type): the Periodic rows look at that (src/solver.jl:458)
    double value[6];
    const double *values[6];  // device arrays (one entry per real cell of the side) or nullptr
};

// reference key of a real border cell (src/solver.jl:379-409): dim2 first, then dim1, then dim3; -1 if interior
__device__ __forceinline__ int border_key(const Grid &g, const int c[PB_MAXD])
{
    if (g.N >= 2) {
        if (c[1] == 0) return PB200_LEFT;
        if (c[1] == g.nc[1] - 1) return PB200_RIGHT;
    }
    if (c[0] == 0) return PB200_BOTTOM;
    if (c[0] == g.nc[0] - 1) return PB200_TOP;
    if (g.N >= 3) {
        if (c[2] == 0) return PB200_BACKWARD;
        if (c[2] == g.nc[2] - 1) return PB200_FORWARD;
    }
    return -1;
}
// index of a border cell inside its side array (other dims, x fastest)
__device__ __forceinline__ int64_t side_index(const Grid &g, int key, const int c[PB_MAXD])
{
    const int dim = (key == PB200_LEFT || key == PB200_RIGHT) ? 1 : (key == PB200_BOTTOM || key == PB200_TOP) ? 0 : 2;
    int64_t idx = 0, str = 1;
    for (int d = 0; d < g.N; ++d) {
        if (d == dim) continue;
        idx += (int64_t)c[d] * str;
        str *= g.nc[d];
    }
    return idx;
}

// Is the bulk row of the real cell c replaced by a row that pins its value, and to what?  (apply_boundary_condition_fast!, src/solver.jl:450-499)
//   Dirichlet: x = value.   Periodic (only if the opposite key is present): x_row - x_partner = 0 with the partner found by
//   find_corresponding_cell_optimized (src/solver.jl:506-530): for a LOW side it is the PAD cell of the opposite end, which the zero-row
//   trimming removes => x_row = 0; for a HIGH side it is the first real cell of that dimension, itself a border cell: pinned => same value.
//   returns 0 not pinned, 1 pinned (val), -1 periodic row whose partner is a free unknown (a genuine coupling row: not supported)
__device__ __forceinline__ int border_pinned(const Grid &g, const BorderDev &bd, const int c[PB_MAXD], double &val, int depth = 0)
{
    val = 0.0;
    const int key = border_key(g, c);
    if (key < 0) return 0;
    const int kind = bd.kind[key];
    if (kind == PB200_BC_DIRICHLET) { val = bd.values[key] ? bd.values[key][side_index(g, key, c)] : bd.value[key]; return 1; }
    if (kind == PB200_BC_PERIODIC) {
        const int opp = key ^ 1;   // LEFT<->RIGHT, BOTTOM<->TOP, BACKWARD<->FORWARD
        if (!bd.present[opp]) return 0;
        const bool hi = key & 1;
        if (!hi) return 1;         // partner = pad cell, removed => 0
        if (depth > 0) return -1;
        const int dim = (key == PB200_LEFT || key == PB200_RIGHT) ? 1 : (key == PB200_BOTTOM || key == PB200_TOP) ? 0 : 2;
        int cp[PB_MAXD] = {c[0], c[1], c[2]};
        cp[dim] = 0;
        const int r = border_pinned(g, bd, cp, val, 1);
        return r == 1 ? 1 : -1;
    }
    return 0;
}

struct SysParams {
    int phase_type, time_type, ifc_kind;
    double alpha, beta;            // mono
    double a1, a2, b1, b2;         // diph
};

// true if the bulk row of cell l (phase ph) has a non-zero coefficient on any T_gamma (H columns) or on a border-Dirichlet neighbour
template <int N>
__device__ __forceinline__ bool row_couples_to_known(const PhaseDev &ph, const Grid &g, int64_t l, const int c[PB_MAXD], const BorderDev &bd)
{
    bool k = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int64_t s = g.stride[d];
        const int id = c[d], last = g.pd[d] - 1;
        const double al = ph.A[d][l], bl = ph.B[d][l];
        k = k || (al != bl);
        if (id > 0) k = k || (al != ph.B[d][l - s]);
        if (id < last) { const double au = ph.A[d][l + s]; k = k || (au != bl) || (au != ph.B[d][l + s]); }
        // border-Dirichlet neighbours
        for (int sg = -1; sg <= 1; sg += 2) {
            int cn[PB_MAXD] = {c[0], c[1], c[2]};
            cn[d] += sg;
            if (cn[d] < 0 || cn[d] >= g.nc[d]) continue;
            bool real = true;
            for (int e = 0; e < N; ++e) real = real && (cn[e] < g.nc[e]);
            if (!real) continue;
            double vv;
            k = k || (border_pinned(g, bd, cn, vv) == 1);
            if (N == 1) { const int kn = border_key(g, cn); k = k || (kn >= 0 && bd.kind[kn] == PB200_BC_NEUMANN); }
        }
    }
    return k;
}

// masks + pinned border values.  mono: phase 0 only.  (BC_border_mono!/diph!, remove_zero_rows_cols!)
template <int N>
__global__ void k_build_masks(Grid g, PhaseDev p1, PhaseDev p2, const double *__restrict__ ct1, const double *__restrict__ ct2, SysParams sp,
                              BorderDev bd, unsigned char *__restrict__ m1, unsigned char *__restrict__ m2, double *__restrict__ ufix1,
                              double *__restrict__ ufix2, int *__restrict__ err)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        bool real = true;
        for (int d = 0; d < N; ++d) real = real && (c[d] < g.nc[d]);
        double bval = 0.0;
        const int pin = real ? border_pinned(g, bd, c, bval) : 0;
        if (pin < 0) *err = 1;
        const bool dirichlet = pin == 1;   // the row pins the value (Dirichlet, or a Periodic row that resolves to a known value)
        bool slave = false;                // 1-D Neumann border row
        if (N == 1 && real && pin == 0) {
            const int kn = border_key(g, c);
            if (kn >= 0 && bd.kind[kn] == PB200_BC_NEUMANN) { slave = true; bval = bd.value[kn] * g.h[0]; }
        }
        const bool unsteady = sp.time_type == PB200_UNSTEADY;
        double GG, HH;
        bool rowG, hrow1, hrow2 = false;
        phase_diag<N>(p1, g, l, c, GG, HH, rowG, hrow1);
        unsigned char b1 = 0, b2 = 0;
        bool kept1 = (unsteady && p1.V[l] != 0.0) || rowG || conv_keeps<N>(p1, g, l, c);
        if (sp.phase_type == PB200_MONO) {
            if (dirichlet) { b1 |= MB_FIXED; ufix1[l] = bval; }
            else if (slave) { b1 |= MB_SLAVE; ufix1[l] = bval; }
            else { if (kept1) b1 |= MB_FREE; ufix1[l] = 0.0; }
            bool keptI = (sp.beta != 0.0 && hrow1) || (sp.alpha != 0.0 && p1.Gam[l] != 0.0);
            if (keptI) b1 |= (sp.beta != 0.0 ? MB_IFREE : MB_IKNOWN);
            if ((b1 & MB_FREE) && row_couples_to_known<N>(p1, g, l, c, bd)) b1 |= MB_KNBR;
            m1[l] = b1;
        } else {
            bool rowG2;
            phase_diag<N>(p2, g, l, c, GG, HH, rowG2, hrow2);
            bool kept2 = (unsteady && p2.V[l] != 0.0) || rowG2 || conv_keeps<N>(p2, g, l, c);
            bool d1 = dirichlet && ct1[l] != 0.0, d2 = dirichlet && ct2[l] != 0.0;   // src/solver.jl:573-576
            const bool s1 = slave && ct1[l] != 0.0, s2 = slave && ct2[l] != 0.0;
            if (d1) { b1 |= MB_FIXED; ufix1[l] = bval; } else if (s1) { b1 |= MB_SLAVE; ufix1[l] = bval; } else { if (kept1) b1 |= MB_FREE; ufix1[l] = 0.0; }
            if (d2) { b2 |= MB_FIXED; ufix2[l] = bval; } else if (s2) { b2 |= MB_SLAVE; ufix2[l] = bval; } else { if (kept2) b2 |= MB_FREE; ufix2[l] = 0.0; }
            bool keptI = (sp.b1 != 0.0 && hrow1) || (sp.b2 != 0.0 && hrow2);
            if (keptI) b2 |= MB_IFREE;   // T_gamma2 is the interface unknown; T_gamma1 = (g + a2 T_gamma2) / a1 everywhere
            if ((b1 & MB_FREE) && row_couples_to_known<N>(p1, g, l, c, bd)) b1 |= MB_KNBR;
            if ((b2 & MB_FREE) && row_couples_to_known<N>(p2, g, l, c, bd)) b2 |= MB_KNBR;
            m1[l] = b1; m2[l] = b2;
        }
    }
}

// Border VALUES changed, kinds did not (time-dependent Dirichlet data, src/solver/diffusion.jl:291-293 calls BC_border_mono! every step):
// only the pinned values are refreshed -- the masks, the folded system and the captured graphs stay valid.
__global__ void k_refresh_ufix(Grid g, BorderDev bd, const unsigned char *__restrict__ m1, const unsigned char *__restrict__ m2, double *__restrict__ ufix1,
                               double *__restrict__ ufix2)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = t + g.plane;
        const unsigned char a = m1[l], b = m2 ? m2[l] : 0;
        if (!((a | b) & (MB_FIXED | MB_SLAVE))) continue;
        int c[PB_MAXD];
        cell_coords(g, t, c);
        double bval = 0.0;
        if (border_pinned(g, bd, c, bval) != 1) {   // 1-D Neumann row
            const int kn = border_key(g, c);
            bval = kn >= 0 ? bd.value[kn] * g.h[0] : 0.0;
        }
        if (a & (MB_FIXED | MB_SLAVE)) ufix1[l] = bval;
        if (b & (MB_FIXED | MB_SLAVE)) ufix2[l] = bval;
    }
}

// ------------------------------------------------------------------------------------------------------------
// general block-row kernels
// ------------------------------------------------------------------------------------------------------------
struct ApplyCoef {
    double cV;    // coefficient of V (1 unsteady, 0 steady)
    double c;     // theta * dt (1 for steady): multiplies D G' W! (...)
    double c2;    // mono interface-row scale (dt/2 for CN, else 1)
    int sym;      // 1: rows scaled to make the mono system symmetric (bulk rows / D_i, interface rows * c/(c2 beta))
    int masked;   // 1: outputs restricted to free rows (Krylov operator); 0: all rows (known-part / explicit products)
};

__device__ __forceinline__ double D_at(const PhaseDev &p, int64_t l) { return p.Darr ? p.Darr[l] : p.Dc; }

// mono: y_b = cV V u + c D Rb ; y_i = c2 (beta Ri + alpha Gamma gam)
template <int N>
__global__ void k_apply_mono(Grid g, PhaseDev p, SysParams sp, ApplyCoef ac, const unsigned char *__restrict__ m, const double *__restrict__ u,
                             GamSpec gs, double *__restrict__ yb, double *__restrict__ yi)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        const unsigned char mb = m[l];
        const bool wb = !ac.masked || (mb & MB_FREE), wi = yi && (!ac.masked || (mb & MB_IFREE));
        double Rb = 0.0, Ri = 0.0;
        if (wb || wi) phase_rows<N>(p, g, l, c, u, gs, Rb, Ri);
        const double D = D_at(p, l);
        if (wb) {
            double v = ac.cV * p.V[l] * (u ? u[l] : 0.0) + ac.c * (D * Rb + conv_row<N>(p, g, l, c, u, gs));
            if (ac.sym) v /= D;
            yb[l] = v;
        } else yb[l] = 0.0;
        if (yi) {
            if (wi) {
                double v = ac.c2 * (sp.beta * Ri + sp.alpha * p.Gam[l] * gam_at(gs, l));
                if (ac.sym) v *= ac.c / (ac.c2 * sp.beta);
                yi[l] = v;
            } else yi[l] = 0.0;
        }
    }
}

// diph: y1 = cV V1 u1 + c D1 Rb1(u1, g1) ; y2 = cV V2 u2 + c D2 Rb2(u2, g2) ; yw = b1 Ri1 + b2 Ri2
template <int N>
__global__ void k_apply_diph(Grid g, PhaseDev p1, PhaseDev p2, SysParams sp, ApplyCoef ac, const unsigned char *__restrict__ m1,
                             const unsigned char *__restrict__ m2, const double *__restrict__ u1, GamSpec g1, const double *__restrict__ u2, GamSpec g2,
                             double *__restrict__ y1, double *__restrict__ y2, double *__restrict__ yw)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        const unsigned char a = m1[l], b = m2[l];
        const bool w1 = !ac.masked || (a & MB_FREE), w2 = !ac.masked || (b & MB_FREE), ww = !ac.masked || (b & MB_IFREE);
        double Rb1 = 0.0, Ri1 = 0.0, Rb2 = 0.0, Ri2 = 0.0;
        if (w1 || ww) phase_rows<N>(p1, g, l, c, u1, g1, Rb1, Ri1);
        if (w2 || ww) phase_rows<N>(p2, g, l, c, u2, g2, Rb2, Ri2);
        y1[l] = w1 ? ac.cV * p1.V[l] * (u1 ? u1[l] : 0.0) + ac.c * (D_at(p1, l) * Rb1 + conv_row<N>(p1, g, l, c, u1, g1)) : 0.0;
        y2[l] = w2 ? ac.cV * p2.V[l] * (u2 ? u2[l] : 0.0) + ac.c * (D_at(p2, l) * Rb2 + conv_row<N>(p2, g, l, c, u2, g2)) : 0.0;
        yw[l] = ww ? sp.b1 * Ri1 + sp.b2 * Ri2 : 0.0;
    }
}

// Jacobi diagonals of the (masked) Krylov operators; entries of non-free rows are set to 1
template <int N>
__global__ void k_diag_mono(Grid g, PhaseDev p, SysParams sp, ApplyCoef ac, const unsigned char *__restrict__ m, double *__restrict__ db,
                            double *__restrict__ di)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        double GG, HH; bool r, h;
        phase_diag<N>(p, g, l, c, GG, HH, r, h);
        const double D = D_at(p, l);
        double vb = ac.cV * p.V[l] + ac.c * (D * GG + conv_diag<N>(p, g, l, c));
        if (vb == 0.0) vb = 1.0;       // (advection: a kept solid cell may have an empty diagonal; Jacobi then leaves the row alone)
        if (ac.sym) vb /= D;
        db[l] = (m[l] & MB_FREE) ? vb : 1.0;
        if (di) {
            double vi = ac.c2 * (sp.beta * HH + sp.alpha * p.Gam[l]);
            if (ac.sym) vi *= ac.c / (ac.c2 * sp.beta);
            di[l] = (m[l] & MB_IFREE) ? vi : 1.0;
        }
    }
}
template <int N>
__global__ void k_diag_diph(Grid g, PhaseDev p1, PhaseDev p2, SysParams sp, ApplyCoef ac, const unsigned char *__restrict__ m1,
                            const unsigned char *__restrict__ m2, double *__restrict__ d1, double *__restrict__ d2, double *__restrict__ dw)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        double GG1, HH1, GG2, HH2; bool r, h;
        phase_diag<N>(p1, g, l, c, GG1, HH1, r, h);
        phase_diag<N>(p2, g, l, c, GG2, HH2, r, h);
        d1[l] = (m1[l] & MB_FREE) ? ac.cV * p1.V[l] + ac.c * (D_at(p1, l) * GG1 + conv_diag<N>(p1, g, l, c)) : 1.0;
        d2[l] = (m2[l] & MB_FREE) ? ac.cV * p2.V[l] + ac.c * (D_at(p2, l) * GG2 + conv_diag<N>(p2, g, l, c)) : 1.0;
        dw[l] = (m2[l] & MB_IFREE) ? sp.b1 * (sp.a2 / sp.a1) * HH1 + sp.b2 * HH2 : 1.0;
    }
}

// grad = W! (G p_omega + H p_gamma)  -- one direction per blockIdx.y  (src/operators.jl:20-23)
template <int N>
__global__ void k_grad(Grid g, PhaseDev p, const double *__restrict__ u, const double *__restrict__ gam, double *__restrict__ out /* N * nloc */)
{
    const int d = blockIdx.y;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane, s = g.stride[d];
        const double bi = p.B[d][l], ai = p.A[d][l];
        const double ei = c[d] < g.pd[d] - 1 ? 1.0 : 0.0;
        double q = ei * (bi * u[l] + (ai - bi) * gam[l]);
        if (c[d] > 0) { const double bm = p.B[d][l - s]; q -= bm * u[l - s] + (ai - bm) * gam[l - s]; }
        out[(int64_t)d * g.nloc + l] = p.Wd[d][l] * q;
    }
}
// div = -(G'+H') q_omega + H' q_gamma  (src/operators.jl:30-34); q_* are N * nloc face fields
template <int N>
__global__ void k_div(Grid g, PhaseDev p, const double *__restrict__ qo, const double *__restrict__ qg, double *__restrict__ out)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
        double acc = 0.0;
        for (int d = 0; d < N; ++d) {
            const int64_t s = g.stride[d], o = (int64_t)d * g.nloc;
            const int last = g.pd[d] - 1;
            const double bi = p.B[d][l], ai = p.A[d][l];
            const double ei = c[d] < last ? 1.0 : 0.0;
            const double hm = ei * (ai - bi);
            double hp = 0.0, qoU = 0.0, qgU = 0.0;
            if (c[d] < last) { hp = p.A[d][l + s] - bi; qoU = qo[o + l + s]; qgU = qg[o + l + s]; }
            const double GTo = bi * (ei * qo[o + l] - qoU);
            const double HTo = hm * qo[o + l] - hp * qoU;
            const double HTg = hm * qg[o + l] - hp * qgU;
            acc += -(GTo + HTo) + HTg;
        }
        out[l] = acc;
    }
}

// W! from W (1/W, 1.0 where W == 0) -- src/operators.jl:145-152
__global__ void k_wdag(int64_t n, const double *__restrict__ W, double *__restrict__ Wd)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double w = W[i];
        Wd[i] = w != 0.0 ? 1.0 / w : 1.0;
    }
}

// 1-D Neumann border rows: the eliminated unknown follows its neighbour (bottom cell <- cell 1, top cell <- cell n-2); restore != 0 puts the
// zero back that Krylov vectors carry on non-free entries
__global__ void k_slave_copy(Grid g, const unsigned char *__restrict__ m, double *__restrict__ u, int restore)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = t + g.plane;
        if (!(m[l] & MB_SLAVE)) continue;
        const int c0 = (int)t + g.k0;
        u[l] = restore ? 0.0 : (c0 == 0 ? u[l + 1] : u[l - 1]);
    }
}
