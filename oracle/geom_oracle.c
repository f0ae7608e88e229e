/*
 * geom_oracle.c -- CPU oracle for the cut-cell geometric moments of `Capacity(levelset, mesh)`.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg through oracle/geom.py.  The product (penguin.jl_b200/) never touches it.
 *
 * What it restates.  /root/reference/src/capacity.jl:81-123 (`VOFI`) obtains every moment from
 * third-party code that is NOT in /root/reference:
 *     CartesianGeometry.jl 0.1.1 (git master, tree 954775b7..., Manifest.toml:223-229)
 *       -> Vofinit.jl 0.1.0 (Manifest.toml:1995-1999) -> libvofi_jll 2.0.0+0 (VOFI 2.0, C),
 *     and C_gamma from ImplicitIntegration.jl 0.1.2 (src/capacity.jl:137-197).
 * VOFI's published algorithm (Bna et al., CPC 2016; Chierici et al., CPC 2022) integrates the *height
 * function* of the implicit surface with Gauss-Legendre rules on sub-intervals split at the kinks.
 * This file restates exactly that idea for the GPU-evaluable level sets of this build (unions of
 * disjoint balls |x-c|-r, axis-aligned half-spaces, and their sign flips): the moments of
 * fluid-in-a-box are nested 1-D integrals of the clipped chord (the height), x outermost, each level
 * split at every event abscissa (tangencies, corner crossings, poles) and integrated by an adaptive
 * Gauss-Kronrod 7/15 rule in a smooth-step variable that removes the square-root end-point
 * singularities.  Meaning/layout of each array follows the in-tree statements of the same quantities:
 *     V, C_omega, cell_types : src/front_tracking.jl:814-897, src/capacity.jl:264-300
 *     A_d (lower face of cell i, faces i = 1..n_d+1) : src/front_tracking.jl:908-1111
 *     W_d (between centroids of i-1 and i, i = 2..n_d), B_d (section through the centroid)
 *                              : src/front_tracking.jl:1124-1326, src/front_tracking1D.jl:214-220
 *     Gamma, C_gamma           : src/front_tracking.jl:1334-1427, src/capacity.jl:137-197
 *     padded layout n = prod(n_i+1), x fastest, pad = 0 : src/capacity.jl:90-92,167-175
 *     geometry grid = mesh.nodes = x0 + (j+1/2) h       : src/mesh.jl:50
 *
 * PARITY UNPINNED at the per-cell 1e-12 / bit-exact level: the reference holds no per-cell golden
 * vectors for this stage and libvofi cannot be run here (no Julia, no libvofi source).  What IS
 * pinned (tests/test_oracle_pins.py, tests/test_geom_oracle.py): the reference's own asserts
 * (area / perimeter / volume vs analytic, `cut <=> Gamma > 0`, centroid-on-circle; test/capacity_test.jl),
 * the closed-form invariants (sum V, sum Gamma, divergence theorem per cell), and an independent
 * high-precision (mpmath) evaluation of sample cells committed under tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC geom_oracle.c -o _build/libgeom_oracle.so -lm
 * (-ffp-contract=off keeps the classification arithmetic plain IEEE mul/add, which the CUDA kernels
 *  reproduce with __dmul_rn/__dadd_rn so that cell_types can be compared bit for bit.)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 3
#define MAXV 4 /* measure + up to 3 first moments */

/* ------------------------------------------------------------------------------------------ */
/* adaptive Gauss-Kronrod 7/15 on a sub-interval [a,b], vector valued, smooth-step variable     */
/* ------------------------------------------------------------------------------------------ */
static const double XGK[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                              0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                              0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                              0.207784955007898467600689403773245, 0.0};
static const double WGK[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                              0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                              0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                              0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
static const double WG[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                             0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

typedef void (*integrand_fn)(double x, void *ctx, double *vals);

static void gk15_panel(integrand_fn f, void *ctx, int nv, double a, double b, double s0, double s1, double *K, double *G)
{
    double hs = 0.5 * (s1 - s0), ms = 0.5 * (s1 + s0), v[MAXV];
    for (int q = 0; q < nv; ++q) K[q] = G[q] = 0.0;
    for (int j = 0; j < 15; ++j) {
        int k = j < 8 ? j : 14 - j;
        double t = j < 8 ? -XGK[k] : XGK[k];
        double s = ms + hs * t;
        double x = a + (b - a) * s * s * (3.0 - 2.0 * s);
        double jac = 6.0 * (b - a) * s * (1.0 - s) * hs;
        f(x, ctx, v);
        for (int q = 0; q < nv; ++q) {
            K[q] += WGK[k] * jac * v[q];
            if (k & 1) G[q] += WG[k / 2] * jac * v[q];
        }
    }
}

/* integrate f over [a,b]; scale[q] = magnitude used for the absolute tolerance of component q */
static void integrate_sub(integrand_fn f, void *ctx, int nv, double a, double b, const double *scale, double *out)
{
    if (!(b > a)) return;
    double st0[64], st1[64];
    int sp = 0;
    st0[0] = 0.0; st1[0] = 1.0; sp = 1;
    while (sp > 0) {
        --sp;
        double s0 = st0[sp], s1 = st1[sp], K[MAXV], G[MAXV];
        gk15_panel(f, ctx, nv, a, b, s0, s1, K, G);
        int ok = 1;
        for (int q = 0; q < nv; ++q)
            if (fabs(K[q] - G[q]) > 2e-13 * scale[q] * (s1 - s0) + 1e-300) ok = 0;
        if (ok && (s1 - s0) >= 1.0 / 8192.0 && sp <= 58) {
            /* second opinion before a panel is accepted: the Kronrod sum of the two halves.  |K - G| alone passes panels whose integrand
             * has a square-root singularity just OUTSIDE the interval (near-tangent sections): tests/host_harness/geom_primitives.cu
             * found 9 of 200 000 random disc/rectangle cases off by 1e-12 .. 5e-11 of the cell area that way (arbitrated with mpmath). */
            double m = 0.5 * (s0 + s1), K1[MAXV], G1[MAXV], K2[MAXV], G2[MAXV];
            gk15_panel(f, ctx, nv, a, b, s0, m, K1, G1);
            gk15_panel(f, ctx, nv, a, b, m, s1, K2, G2);
            int agree = 1;
            for (int q = 0; q < nv; ++q)
                if (fabs(K1[q] + K2[q] - K[q]) > 2e-14 * scale[q] * (s1 - s0) + 1e-300) agree = 0;
            if (agree) { for (int q = 0; q < nv; ++q) out[q] += K1[q] + K2[q]; continue; }
            ok = 0;
        }
        if (ok || (s1 - s0) < 1.0 / 8192.0 || sp > 60) {   /* depth cap: the integrand is smooth in s */
            for (int q = 0; q < nv; ++q) out[q] += K[q];
        } else {
            double m = 0.5 * (s0 + s1);
            st0[sp] = s0; st1[sp] = m; ++sp;
            st0[sp] = m; st1[sp] = s1; ++sp;
        }
    }
}

static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

/* integrate over [a,b] split at the event abscissae ev[0..ne) */
static void integrate_events(integrand_fn f, void *ctx, int nv, double a, double b, double *ev, int ne, const double *scale, double *out)
{
    if (!(b > a)) return;
    double pts[80];
    int np = 0;
    pts[np++] = a;
    for (int i = 0; i < ne && np < 78; ++i)
        if (ev[i] > a && ev[i] < b) pts[np++] = ev[i];
    pts[np++] = b;
    qsort(pts, np, sizeof(double), cmp_d);
    for (int i = 0; i + 1 < np; ++i)
        if (pts[i + 1] - pts[i] > 1e-15 * (fabs(a) + fabs(b) + (b - a))) integrate_sub(f, ctx, nv, pts[i], pts[i + 1], scale, out);
}

/* event abscissae along axis 0 for a ball (centre c, radius^2 R2) against the (m-1)-box lo[1..],hi[1..] */
static int ball_events(int m, const double *c, double R2, const double *lo, const double *hi, double *ev)
{
    int ne = 0;
    double R = sqrt(R2);
    ev[ne++] = c[0] - R; ev[ne++] = c[0] + R;
    int nd = m - 1;
    /* every non-empty subset of the remaining dims, every lo/hi choice */
    for (int mask = 1; mask < (1 << nd); ++mask) {
        int dims[MAXD], k = 0;
        for (int e = 0; e < nd; ++e) if (mask & (1 << e)) dims[k++] = e + 1;
        for (int ch = 0; ch < (1 << k); ++ch) {
            double d2 = 0.0;
            for (int q = 0; q < k; ++q) {
                double dl = ((ch >> q) & 1 ? hi[dims[q]] : lo[dims[q]]) - c[dims[q]];
                d2 += dl * dl;
            }
            if (d2 < R2) { double s = sqrt(R2 - d2); ev[ne++] = c[0] - s; ev[ne++] = c[0] + s; }
        }
    }
    return ne;
}

/* ------------------------------------------------------------------------------------------ */
/* ball ∩ box: measure and first moments about `mid` (m = 0..3 dims)                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int m; const double *c, *lo, *hi, *mid; double R2; } bbctx;
static void bb_moments(int m, const double *c, double R2, const double *lo, const double *hi, const double *mid, double *out);

static void bb_integrand(double x, void *vctx, double *vals)
{
    bbctx *k = (bbctx *)vctx;
    double sub[MAXV] = {0, 0, 0, 0};
    double dx = x - k->c[0];
    bb_moments(k->m - 1, k->c + 1, k->R2 - dx * dx, k->lo + 1, k->hi + 1, k->mid + 1, sub);
    vals[0] = sub[0];
    vals[1] = (x - k->mid[0]) * sub[0];
    for (int q = 1; q < k->m; ++q) vals[1 + q] = sub[q];
}

static void bb_moments(int m, const double *c, double R2, const double *lo, const double *hi, const double *mid, double *out)
{
    for (int q = 0; q <= m; ++q) out[q] = 0.0;
    if (m == 0) { out[0] = R2 > 0.0 ? 1.0 : 0.0; return; }
    if (!(R2 > 0.0)) return;
    double R = sqrt(R2);
    if (m == 1) {
        double a = fmax(lo[0], c[0] - R), b = fmin(hi[0], c[0] + R);
        if (b > a) { out[0] = b - a; out[1] = 0.5 * ((b - mid[0]) * (b - mid[0]) - (a - mid[0]) * (a - mid[0])); }
        return;
    }
    double a = fmax(lo[0], c[0] - R), b = fmin(hi[0], c[0] + R);
    if (!(b > a)) return;
    double ev[64];
    int ne = ball_events(m, c, R2, lo, hi, ev);
    bbctx k = {m, c, lo, hi, mid, R2};
    double scale[MAXV], cross = 1.0;
    for (int e = 1; e < m; ++e) cross *= hi[e] - lo[e];
    scale[0] = cross;
    scale[1] = cross * (hi[0] - lo[0]);
    for (int q = 1; q < m; ++q) scale[1 + q] = cross * (hi[q] - lo[q]);
    integrate_events(bb_integrand, &k, m + 1, a, b, ev, ne, scale, out);
}

/* ------------------------------------------------------------------------------------------ */
/* sphere ∩ box: surface measure and first moments about `mid`                                  */
/*   m = 1: the two points c +- R (counting measure)                                            */
/*   m = 2: exact, by the angular intervals of the circle that lie inside the rectangle         */
/*   m = 3: hat-box form dS = (R / rho(x)) ds dx, i.e. S = R * int dphi(x) dx, outer GK in x     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const double *c, *lo, *hi, *mid; double R; } sbctx;
static void sb_area(int m, const double *c, double R, const double *lo, const double *hi, const double *mid, double *out);

static void circle_rect_arcs(const double *c, double rho, const double *lo, const double *hi, const double *mid, double *out)
{
    const double TWO_PI = 6.283185307179586476925286766559;
    double ang[10];
    int na = 0;
    for (int side = 0; side < 2; ++side) {
        double u = ((side ? hi[0] : lo[0]) - c[0]) / rho;
        if (fabs(u) < 1.0) { double a = acos(u); ang[na++] = a; ang[na++] = TWO_PI - a; }
        double v = ((side ? hi[1] : lo[1]) - c[1]) / rho;
        if (fabs(v) < 1.0) { double a = asin(v); ang[na++] = a < 0 ? a + TWO_PI : a; ang[na++] = 3.1415926535897932384626433832795 - a; }
    }
    out[0] = out[1] = out[2] = 0.0;
    if (na == 0) {
        double py = c[0] + rho, pz = c[1];
        if (py > lo[0] && py < hi[0] && pz > lo[1] && pz < hi[1] && c[0] - rho > lo[0] && c[1] + rho < hi[1] && c[1] - rho > lo[1]) {
            out[0] = TWO_PI * rho; out[1] = out[0] * (c[0] - mid[0]); out[2] = out[0] * (c[1] - mid[1]);
        }
        return;
    }
    qsort(ang, na, sizeof(double), cmp_d);
    ang[na] = ang[0] + TWO_PI;
    for (int i = 0; i < na; ++i) {
        double a0 = ang[i], a1 = ang[i + 1];
        if (!(a1 > a0)) continue;
        double am = 0.5 * (a0 + a1), py = c[0] + rho * cos(am), pz = c[1] + rho * sin(am);
        if (py >= lo[0] && py <= hi[0] && pz >= lo[1] && pz <= hi[1]) {   /* closed: an arc tangent to a face at its mid-point is inside */
            double dphi = a1 - a0;
            out[0] += rho * dphi;
            out[1] += rho * ((c[0] - mid[0]) * dphi + rho * (sin(a1) - sin(a0)));
            out[2] += rho * ((c[1] - mid[1]) * dphi - rho * (cos(a1) - cos(a0)));
        }
    }
}

static void sb_integrand(double x, void *vctx, double *vals)
{
    sbctx *k = (sbctx *)vctx;
    double sub[3];
    double dx = x - k->c[0], r2 = k->R * k->R - dx * dx;
    vals[0] = vals[1] = vals[2] = vals[3] = 0.0;
    if (!(r2 > 0.0)) return;
    double rho = sqrt(r2), fac = k->R / rho;
    circle_rect_arcs(k->c + 1, rho, k->lo + 1, k->hi + 1, k->mid + 1, sub);
    vals[0] = fac * sub[0];
    vals[1] = fac * (x - k->mid[0]) * sub[0];
    vals[2] = fac * sub[1];
    vals[3] = fac * sub[2];
}

static void sb_area(int m, const double *c, double R, const double *lo, const double *hi, const double *mid, double *out)
{
    for (int q = 0; q <= m; ++q) out[q] = 0.0;
    if (!(R > 0.0)) return;
    if (m == 1) {
        double p0 = c[0] - R, p1 = c[0] + R;
        if (p0 >= lo[0] && p0 < hi[0]) { out[0] += 1.0; out[1] += p0 - mid[0]; }
        if (p1 >= lo[0] && p1 < hi[0]) { out[0] += 1.0; out[1] += p1 - mid[0]; }
        return;
    }
    if (m == 2) { circle_rect_arcs(c, R, lo, hi, mid, out); return; }
    double a = fmax(lo[0], c[0] - R), b = fmin(hi[0], c[0] + R);
    if (!(b > a)) return;
    double ev[64];
    int ne = ball_events(m, c, R * R, lo, hi, ev);
    sbctx k = {c, lo, hi, mid, R};
    double scale[MAXV], cross = 1.0, hmax = 0.0;
    for (int e = 1; e < m; ++e) cross *= hi[e] - lo[e];
    for (int e = 0; e < m; ++e) hmax = fmax(hmax, hi[e] - lo[e]);
    scale[0] = cross;
    for (int q = 0; q < m; ++q) scale[1 + q] = cross * hmax;
    integrate_events(sb_integrand, &k, m + 1, a, b, ev, ne, scale, out);
}

/* ------------------------------------------------------------------------------------------ */
/* level-set description                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int N;
    int kind;          /* 0: union of disjoint balls, 1: axis-aligned half-space {x_hd < hc} */
    int nb;
    const double *c;   /* nb*N */
    const double *r;   /* nb */
    int inside;        /* 1: fluid = {phi < 0} = inside balls / x < hc ; 0: complement */
    int hd; double hc;
} shape;

/* classification of a box against the "in" set (before the sign flip): 1 full-in, 0 empty, -1 cut */
static int classify_in(const shape *s, int m, const int *dims, const double *lo, const double *hi, int fixd, double fixv)
{
    /* box spans dims[0..m) with bounds lo/hi (indexed by position); optionally one fixed coordinate */
    if (s->kind == 1) {
        if (fixd == s->hd) return fixv < s->hc ? 1 : 0;
        for (int q = 0; q < m; ++q)
            if (dims[q] == s->hd) { if (hi[q] <= s->hc) return 1; if (lo[q] >= s->hc) return 0; return -1; }
        return 0;
    }
    int res = 0;
    for (int b = 0; b < s->nb; ++b) {
        const double *c = s->c + (size_t)b * s->N;
        double R2 = s->r[b] * s->r[b], dmin2 = 0.0, dmax2 = 0.0;
        for (int q = 0; q < m; ++q) {
            double cl = c[dims[q]];
            double a = lo[q] - cl, bb = cl - hi[q];
            double dn = fmax(fmax(a, bb), 0.0);
            double dx = fmax(fabs(a), fabs(bb));
            dmin2 = dmin2 + dn * dn;
            dmax2 = dmax2 + dx * dx;
        }
        if (fixd >= 0) { double f = fixv - c[fixd]; dmin2 = dmin2 + f * f; dmax2 = dmax2 + f * f; }
        if (dmax2 <= R2) return 1;
        if (dmin2 < R2) res = -1;
    }
    return res;
}

/* measure + first moments (about mid) of the "in" set inside a box spanning dims[0..m), optional fixed coord */
static void in_moments(const shape *s, int m, const int *dims, const double *lo, const double *hi, const double *mid,
                       int fixd, double fixv, double *out)
{
    for (int q = 0; q <= m; ++q) out[q] = 0.0;
    if (s->kind == 1) {
        double l2[MAXD], h2[MAXD];
        double meas = 1.0;
        int hit = 0;
        for (int q = 0; q < m; ++q) { l2[q] = lo[q]; h2[q] = hi[q]; if (dims[q] == s->hd) { h2[q] = fmin(hi[q], s->hc); hit = 1; } }
        if (fixd == s->hd) { if (!(fixv < s->hc)) return; }
        else if (!hit) return;
        for (int q = 0; q < m; ++q) { if (!(h2[q] > l2[q])) return; meas *= h2[q] - l2[q]; }
        out[0] = meas;
        for (int q = 0; q < m; ++q) out[1 + q] = meas * (0.5 * (l2[q] + h2[q]) - mid[q]);
        return;
    }
    for (int b = 0; b < s->nb; ++b) {
        const double *cb = s->c + (size_t)b * s->N;
        double c[MAXD], R2 = s->r[b] * s->r[b], sub[MAXV];
        for (int q = 0; q < m; ++q) c[q] = cb[dims[q]];
        if (fixd >= 0) { double f = fixv - cb[fixd]; R2 -= f * f; }
        bb_moments(m, c, R2, lo, hi, mid, sub);
        for (int q = 0; q <= m; ++q) out[q] += sub[q];
    }
}

static void interface_moments(const shape *s, int m, const double *lo, const double *hi, const double *mid, double *out)
{
    for (int q = 0; q <= m; ++q) out[q] = 0.0;
    if (s->kind == 1) {
        if (!(lo[s->hd] < s->hc && s->hc < hi[s->hd])) return;
        double meas = 1.0;
        for (int q = 0; q < m; ++q) if (q != s->hd) meas *= hi[q] - lo[q];
        out[0] = meas;
        for (int q = 0; q < m; ++q) out[1 + q] = q == s->hd ? meas * (s->hc - mid[q]) : 0.0;
        return;
    }
    for (int b = 0; b < s->nb; ++b) {
        double sub[MAXV];
        sb_area(m, s->c + (size_t)b * s->N, s->r[b], lo, hi, mid, sub);
        for (int q = 0; q <= m; ++q) out[q] += sub[q];
    }
}

/* fluid measure (+ moments) of a box, including the sign flip; also returns the type */
static int fluid_box(const shape *s, int m, const int *dims, const double *lo, const double *hi, int fixd, double fixv,
                     double *meas, double *bary /* may be NULL; size m */)
{
    double full = 1.0, mid[MAXD];
    for (int q = 0; q < m; ++q) { full *= hi[q] - lo[q]; mid[q] = 0.5 * (lo[q] + hi[q]); }
    int t = classify_in(s, m, dims, lo, hi, fixd, fixv);
    if (!s->inside && t >= 0) t = 1 - t;
    if (t == 1) { *meas = full; if (bary) for (int q = 0; q < m; ++q) bary[q] = mid[q]; return 1; }
    if (t == 0) { *meas = 0.0; if (bary) for (int q = 0; q < m; ++q) bary[q] = mid[q]; return 0; }
    double mom[MAXV];
    in_moments(s, m, dims, lo, hi, mid, fixd, fixv, mom);
    if (!s->inside) { mom[0] = full - mom[0]; for (int q = 0; q < m; ++q) mom[1 + q] = -mom[1 + q]; }
    *meas = mom[0];
    if (bary) for (int q = 0; q < m; ++q) bary[q] = mom[0] > 0.0 ? mid[q] + mom[1 + q] / mom[0] : mid[q];
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* public entry: all capacities of src/capacity.jl:81-123 on the padded grid                    */
/* ------------------------------------------------------------------------------------------ */
int pgo_capacity(int N, const int *ncell, const double *x0, const double *L,
                 int kind, int nb, const double *centers, const double *radii, int inside, int hd, double hc,
                 double *V, double *Gamma, double *ctype, double *A, double *B, double *W, double *Com, double *Cga)
{
    if (N < 1 || N > 3) return 1;
    shape s = {N, kind, nb, centers, radii, inside, hd, hc};
    int pd[MAXD] = {1, 1, 1}, nc[MAXD] = {1, 1, 1};
    double h[MAXD] = {1, 1, 1}, *nodes[MAXD];
    size_t n = 1;
    for (int d = 0; d < N; ++d) {
        nc[d] = ncell[d]; pd[d] = ncell[d] + 1; n *= (size_t)pd[d];
        h[d] = L[d] / ncell[d];
        nodes[d] = (double *)malloc(sizeof(double) * (size_t)(pd[d] + 1));
        for (int j = 0; j <= pd[d]; ++j) nodes[d][j] = x0[d] + (j + 0.5) * h[d];   /* src/mesh.jl:50 */
    }
    memset(V, 0, n * sizeof(double)); memset(Gamma, 0, n * sizeof(double)); memset(ctype, 0, n * sizeof(double));
    memset(A, 0, N * n * sizeof(double)); memset(B, 0, N * n * sizeof(double)); memset(W, 0, N * n * sizeof(double));
    memset(Com, 0, N * n * sizeof(double));
    if (Cga) memset(Cga, 0, N * n * sizeof(double));
    int alld[MAXD] = {0, 1, 2};

    /* pass 1: V, barycentre, type, Gamma, C_gamma on real cells */
    for (int k = 0; k < nc[2]; ++k) for (int j = 0; j < nc[1]; ++j) for (int i = 0; i < nc[0]; ++i) {
        int ix[MAXD] = {i, j, k};
        size_t idx = (size_t)i + (size_t)pd[0] * ((size_t)j + (size_t)pd[1] * (size_t)k);
        double lo[MAXD], hi[MAXD], bary[MAXD], meas;
        for (int d = 0; d < N; ++d) { lo[d] = nodes[d][ix[d]]; hi[d] = nodes[d][ix[d] + 1]; }
        int t = fluid_box(&s, N, alld, lo, hi, -1, 0.0, &meas, bary);
        V[idx] = meas; ctype[idx] = (double)t;
        for (int d = 0; d < N; ++d) Com[(size_t)d * n + idx] = bary[d];
        if (t == -1) {
            double mid[MAXD], g[MAXV];
            for (int d = 0; d < N; ++d) mid[d] = 0.5 * (lo[d] + hi[d]);
            interface_moments(&s, N, lo, hi, mid, g);
            Gamma[idx] = g[0];
            if (Cga && g[0] > 0.0) for (int d = 0; d < N; ++d) Cga[(size_t)d * n + idx] = mid[d] + g[1 + d] / g[0];
        }
    }
    /* pass 2: A_d, B_d, W_d */
    for (int d = 0; d < N; ++d) {
        int od[MAXD], m = 0;
        for (int e = 0; e < N; ++e) if (e != d) od[m++] = e;
        for (int k = 0; k < pd[2]; ++k) for (int j = 0; j < pd[1]; ++j) for (int i = 0; i < pd[0]; ++i) {
            int ix[MAXD] = {i, j, k};
            int real_others = 1;
            for (int q = 0; q < m; ++q) if (ix[od[q]] >= nc[od[q]]) real_others = 0;
            if (!real_others) continue;
            size_t idx = (size_t)i + (size_t)pd[0] * ((size_t)j + (size_t)pd[1] * (size_t)k);
            size_t str = 1; for (int e = 0; e < d; ++e) str *= (size_t)pd[e];
            double lo[MAXD], hi[MAXD], meas;
            for (int q = 0; q < m; ++q) { lo[q] = nodes[od[q]][ix[od[q]]]; hi[q] = nodes[od[q]][ix[od[q]] + 1]; }
            /* A_d: lower face of cell ix (face index 0..n_d) */
            fluid_box(&s, m, od, lo, hi, d, nodes[d][ix[d]], &meas, NULL);
            A[(size_t)d * n + idx] = meas;
            if (ix[d] < nc[d]) {
                /* B_d: section through the barycentre of this (real) cell */
                double t = ctype[idx], face = 1.0;
                for (int q = 0; q < m; ++q) face *= hi[q] - lo[q];
                if (t == 1.0) B[(size_t)d * n + idx] = face;
                else if (t == 0.0) B[(size_t)d * n + idx] = 0.0;
                else { fluid_box(&s, m, od, lo, hi, d, Com[(size_t)d * n + idx], &meas, NULL); B[(size_t)d * n + idx] = meas; }
                /* W_d: staggered box between the barycentres of cell ix-1 and ix (ix[d] = 1..n_d-1) */
                if (ix[d] >= 1) {
                    double blo[MAXD], bhi[MAXD];
                    for (int e = 0; e < N; ++e) { blo[e] = nodes[e][ix[e]]; bhi[e] = nodes[e][ix[e] + 1]; }
                    blo[d] = Com[(size_t)d * n + idx - str];
                    bhi[d] = Com[(size_t)d * n + idx];
                    fluid_box(&s, N, alld, blo, bhi, -1, 0.0, &meas, NULL);
                    W[(size_t)d * n + idx] = meas;
                }
            }
        }
    }
    for (int d = 0; d < N; ++d) free(nodes[d]);
    return 0;
}

/* exposed primitives (used by the golden-vector checks) */
void pgo_ball_box(int m, const double *c, double R, const double *lo, const double *hi, double *out /* m+1 */)
{
    double mid[MAXD];
    for (int q = 0; q < m; ++q) mid[q] = 0.5 * (lo[q] + hi[q]);
    bb_moments(m, c, R * R, lo, hi, mid, out);
}
void pgo_sphere_box(int m, const double *c, double R, const double *lo, const double *hi, double *out /* m+1 */)
{
    double mid[MAXD];
    for (int q = 0; q < m; ++q) mid[q] = 0.5 * (lo[q] + hi[q]);
    sb_area(m, c, R, lo, hi, mid, out);
}
