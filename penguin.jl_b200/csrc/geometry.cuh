// geometry.cuh -- cut-cell geometric moments on the device: replaces `VOFI(body, mesh)` and `computeInterfaceCentroids`
// (/root/reference/src/capacity.jl:81-123, 137-197; upstream arithmetic in CartesianGeometry.jl -> libvofi).
//
// Level sets are unions of disjoint balls (interval / circle / sphere), axis-aligned half-spaces, and their sign flips.
// Method (independent of the CPU oracle, which nests Gauss-Kronrod on chord heights with x outermost):
//   * disc /\ rectangle is EXACT: the region's boundary is walked once around the rectangle; the area and first moments
//     are a polygon (shoelace, local coordinates) plus circular segments, each segment evaluated about its own chord
//     mid-point with series for small half-angles, so that nothing cancels when R >> h;
//   * ball /\ box integrates that exact section over z (outermost) between all event heights (tangencies, corner
//     crossings, poles) with an adaptive Gauss-Kronrod 7/15 rule in a smooth-step variable; the sphere area uses the
//     hat-box form dS = R dphi dz with the exact arc angles of the section;
//   * classification uses plain IEEE mul/add (__dmul_rn/__dadd_rn, no FMA contraction) in the same order as the oracle,
//     so cell_types can be compared bit for bit.
// Pipeline: k_geom_cells (classify, full/empty fill, exact 1-D/2-D cut cells, 3-D cut list) -> k_geom_cut3d (compact list)
//           -> k_geom_faces (A, B, exact 1-D/2-D W, 3-D W list) -> k_geom_w3d (compact list).
#pragma once
#include "common.cuh"

#define DISPATCH_GEOM_N(N_, ...)                  \
    do {                                     \
        if ((N_) == 1) { constexpr int NN = 1; __VA_ARGS__; } \
        else if ((N_) == 2) { constexpr int NN = 2; __VA_ARGS__; } \
        else { constexpr int NN = 3; __VA_ARGS__; } \
    } while (0)

struct GeomOut { double *V, *Gam, *ct, *A[PB_MAXD], *B[PB_MAXD], *W[PB_MAXD], *Co[PB_MAXD], *Cg[PB_MAXD]; };

struct ShapeDev {
    int kind, nb, inside, hd;
    const double *c;  // nb * N
    const double *r;  // nb
    double hc;
};

#define PB_PI 3.14159265358979323846264338327950288
#define PB_TWO_PI 6.28318530717958647692528676655900577

// ---- small-angle safe segment / arc functions ------------------------------------------------------------------------
// segA(p) = p - sin p cos p ; segM(p) = 2/3 sin^3 p - segA(p) cos p ; arcO(p) = sin p / p - cos p
__host__ __device__ __forceinline__ double seg_area_f(double p, double sp, double cp)
{
    if (p < 0.25) {
        const double x2 = p * p;
        return p * x2 * (2.0 / 3.0 + x2 * (-2.0 / 15.0 + x2 * (4.0 / 315.0 + x2 * (-2.0 / 2835.0 + x2 * (4.0 / 155925.0 + x2 * (-4.0 / 6081075.0 + x2 * (8.0 / 638512875.0)))))));
    }
    return p - sp * cp;
}
__host__ __device__ __forceinline__ double seg_mom_f(double p, double sp, double cp)
{
    if (p < 0.25) {
        const double x2 = p * p;
        return p * x2 * x2 * (2.0 / 15.0 + x2 * (-11.0 / 315.0 + x2 * (17.0 / 3780.0 + x2 * (-461.0 / 1247400.0 + x2 * (8303.0 / 389188800.0 + x2 * (-24911.0 / 27243216000.0))))));
    }
    return (2.0 / 3.0) * sp * sp * sp - (p - sp * cp) * cp;
}
__host__ __device__ __forceinline__ double arc_off_f(double p, double sp, double cp)
{
    if (p < 0.25) {
        const double x2 = p * p;
        return x2 * (1.0 / 3.0 + x2 * (-1.0 / 30.0 + x2 * (1.0 / 840.0 + x2 * (-1.0 / 45360.0 + x2 * (1.0 / 3991680.0 + x2 * (-1.0 / 518918400.0))))));
    }
    return sp / p - cp;
}

// ---- exact disc /\ rectangle -------------------------------------------------------------------------------------------
// local frame: rectangle [-hx,hx] x [-hy,hy], disc centre (X0,Y0), radius rho.
// out[0] area, out[1..2] first moments about the rectangle centre, out[3] sum of arc angles inside the rectangle,
// out[4..5] sum over arcs of theta * (arc centroid - rectangle centre)   (arc LENGTH quantities are rho * these)
__host__ __device__ __noinline__ void disc_rect(double X0, double Y0, double rho, double hx, double hy, double *out)
{
    for (int q = 0; q < 6; ++q) out[q] = 0.0;
    const double r2 = rho * rho;
    const double ax = fabs(X0), ay = fabs(Y0);
    {
        const double fx = ax + hx, fy = ay + hy;
        if (fx * fx + fy * fy <= r2) { out[0] = 4.0 * hx * hy; return; }
        const double nx = fmax(ax - hx, 0.0), ny = fmax(ay - hy, 0.0);
        if (nx * nx + ny * ny >= r2) return;
    }
    double vx[12], vy[12];
    unsigned ex = 0u;  // bit i: vertex i is an exit crossing (the perimeter leaves the disc after it)
    int nv = 0;
    // edge 0: bottom (y=-hy, +x), 1: right (x=hx, +y), 2: top (y=hy, -x), 3: left (x=-hx, -y)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double px = (e == 0 || e == 3) ? -hx : hx, py = (e < 2) ? -hy : hy;
        {
            const double dx = px - X0, dy = py - Y0;
            if (dx * dx + dy * dy < r2) { vx[nv] = px; vy[nv] = py; ++nv; }
        }
        if ((e & 1) == 0) {
            const double dy = py - Y0, q = r2 - dy * dy;
            if (q > 0.0) {
                const double sq = sqrt(q), x1 = X0 - sq, x2 = X0 + sq;
                if (e == 0) {   // +x: x1 entry, x2 exit ; positions in [-hx, hx)
                    if (x1 >= -hx && x1 < hx) { vx[nv] = x1; vy[nv] = py; ++nv; }
                    if (x2 >= -hx && x2 < hx) { vx[nv] = x2; vy[nv] = py; ex |= 1u << nv; ++nv; }
                } else {        // -x: x2 entry, x1 exit ; positions in (-hx, hx]
                    if (x2 > -hx && x2 <= hx) { vx[nv] = x2; vy[nv] = py; ++nv; }
                    if (x1 > -hx && x1 <= hx) { vx[nv] = x1; vy[nv] = py; ex |= 1u << nv; ++nv; }
                }
            }
        } else {
            const double dx = px - X0, q = r2 - dx * dx;
            if (q > 0.0) {
                const double sq = sqrt(q), y1 = Y0 - sq, y2 = Y0 + sq;
                if (e == 1) {   // +y
                    if (y1 >= -hy && y1 < hy) { vx[nv] = px; vy[nv] = y1; ++nv; }
                    if (y2 >= -hy && y2 < hy) { vx[nv] = px; vy[nv] = y2; ex |= 1u << nv; ++nv; }
                } else {        // -y
                    if (y2 > -hy && y2 <= hy) { vx[nv] = px; vy[nv] = y2; ++nv; }
                    if (y1 > -hy && y1 <= hy) { vx[nv] = px; vy[nv] = y1; ex |= 1u << nv; ++nv; }
                }
            }
        }
    }
    if (nv < 2) {
        // no crossing and no corner inside: the disc is entirely inside the rectangle (the empty case was excluded above)
        if (ax < hx && ay < hy) {
            const double a = PB_PI * r2;
            out[0] = a; out[1] = a * X0; out[2] = a * Y0;
            out[3] = PB_TWO_PI; out[4] = PB_TWO_PI * X0; out[5] = PB_TWO_PI * Y0;
        }
        return;
    }
    double area = 0.0, mx = 0.0, my = 0.0, th = 0.0, tx = 0.0, ty = 0.0;
    for (int i = 0; i < nv; ++i) {
        const int j = i + 1 < nv ? i + 1 : 0;
        const double x0 = vx[i], y0 = vy[i], x1 = vx[j], y1 = vy[j];
        const double cr = x0 * y1 - x1 * y0;
        area += 0.5 * cr;
        mx += (x0 + x1) * cr * (1.0 / 6.0);
        my += (y0 + y1) * cr * (1.0 / 6.0);
        if (ex & (1u << i)) {
            // circular arc from P=(x0,y0) to Q=(x1,y1), counter-clockwise about the disc centre
            const double dx = x1 - x0, dy = y1 - y0;
            const double c = sqrt(dx * dx + dy * dy);
            if (c > 0.0) {
                double s = c / (2.0 * rho);
                if (s > 1.0) s = 1.0;
                // half angle p of the arc from its chord: tan p = (c / 2) / d, d = crs / c = signed distance of the disc centre from the
                // chord (negative: major arc), cos p = d / rho.  (asin(c / 2 rho) and sqrt(1 - s^2) lose up to eight digits when the chord
                // is close to a diameter -- a disc centre on or near an edge line, e.g. every z-section of a sphere centred on a grid
                // plane: tests/host_harness/geom_primitives.cu, half-ball volume off by 2e-10 relative.)
                const double crs = (x0 - X0) * (y1 - Y0) - (x1 - X0) * (y0 - Y0);
                const double d = crs / c;
                const double p = atan2(0.5 * c, d);
                double cp = d / rho;
                cp = cp > 1.0 ? 1.0 : (cp < -1.0 ? -1.0 : cp);
                const double nxh = dy / c, nyh = -dx / c;      // unit vector from the chord mid-point towards the arc
                const double Mx = 0.5 * (x0 + x1), My = 0.5 * (y0 + y1);
                const double sa = r2 * seg_area_f(p, s, cp);
                const double sm = r2 * rho * seg_mom_f(p, s, cp);
                area += sa;
                mx += sa * Mx + sm * nxh;
                my += sa * My + sm * nyh;
                const double t2 = 2.0 * p, off = rho * arc_off_f(p, s, cp);
                th += t2;
                tx += t2 * (Mx + off * nxh);
                ty += t2 * (My + off * nyh);
            }
        }
    }
    out[0] = area; out[1] = mx; out[2] = my; out[3] = th; out[4] = tx; out[5] = ty;
}

// chord [c-r, c+r] /\ [lo,hi]: length and first moment about mid
__host__ __device__ __forceinline__ void chord_seg(double c, double r2, double lo, double hi, double mid, double &len, double &mom)
{
    len = 0.0; mom = 0.0;
    if (!(r2 > 0.0)) return;
    const double r = sqrt(r2);
    const double a = fmax(lo, c - r), b = fmin(hi, c + r);
    if (b > a) { len = b - a; mom = 0.5 * ((b - mid) * (b - mid) - (a - mid) * (a - mid)); }
}

// ---- adaptive Gauss-Kronrod 7/15 over z of the exact sections --------------------------------------------------------
__constant__ double c_xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851, 0.864864423359769072789712788640926,
                                0.741531185599394439863864773280788, 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.0};
__constant__ double c_wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
                                0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
                               0.417959183673469387755102040816327};
static const double h_xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851, 0.864864423359769072789712788640926,
                                0.741531185599394439863864773280788, 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.0};
static const double h_wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
                                0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
static const double h_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
                               0.417959183673469387755102040816327};
#ifdef __CUDA_ARCH__
#define GK_X(k) c_xgk[k]
#define GK_WK(k) c_wgk[k]
#define GK_WG(k) c_wg[k]
#else
#define GK_X(k) h_xgk[k]
#define GK_WK(k) h_wgk[k]
#define GK_WG(k) h_wg[k]
#endif

struct BallBox {
    double c[3], R;       // ball
    double mid[3], hw[3]; // box centre and half widths
};

// integrand at height z: mode 0 -> volume moments [V, Mx, My, Mz]; mode 1 -> surface [S, Sx, Sy, Sz] (about the box centre)
__host__ __device__ __forceinline__ void bb_section(const BallBox &b, int mode, double z, double *v)
{
    const double dz = z - b.c[2];
    const double r2 = b.R * b.R - dz * dz;
    v[0] = v[1] = v[2] = v[3] = 0.0;
    if (!(r2 > 0.0)) return;
    double o[6];
    disc_rect(b.c[0] - b.mid[0], b.c[1] - b.mid[1], sqrt(r2), b.hw[0], b.hw[1], o);
    if (mode == 0) { v[0] = o[0]; v[1] = o[1]; v[2] = o[2]; v[3] = (z - b.mid[2]) * o[0]; }
    else { v[0] = b.R * o[3]; v[1] = b.R * o[4]; v[2] = b.R * o[5]; v[3] = b.R * (z - b.mid[2]) * o[3]; }
}

__host__ __device__ __noinline__ void bb_integrate(const BallBox &b, int mode, double *res)
{
    res[0] = res[1] = res[2] = res[3] = 0.0;
    const double zlo = fmax(b.mid[2] - b.hw[2], b.c[2] - b.R), zhi = fmin(b.mid[2] + b.hw[2], b.c[2] + b.R);
    if (!(zhi > zlo)) return;
    // event heights: z = cz +- sqrt(R^2 - d^2) for the distances d of the section centre to the rectangle's edge lines / corners
    double ev[20];
    int ne = 0;
    ev[ne++] = zlo; ev[ne++] = zhi;
    const double R2 = b.R * b.R;
    const double dxl = b.mid[0] - b.hw[0] - b.c[0], dxh = b.mid[0] + b.hw[0] - b.c[0];
    const double dyl = b.mid[1] - b.hw[1] - b.c[1], dyh = b.mid[1] + b.hw[1] - b.c[1];
    const double d2s[8] = {dxl * dxl, dxh * dxh, dyl * dyl, dyh * dyh, dxl * dxl + dyl * dyl, dxl * dxl + dyh * dyh, dxh * dxh + dyl * dyl, dxh * dxh + dyh * dyh};
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (d2s[k] < R2) {
            const double s = sqrt(R2 - d2s[k]);
            const double z1 = b.c[2] - s, z2 = b.c[2] + s;
            if (z1 > zlo && z1 < zhi) ev[ne++] = z1;
            if (z2 > zlo && z2 < zhi) ev[ne++] = z2;
        }
    for (int i = 1; i < ne; ++i) {   // insertion sort
        const double x = ev[i];
        int j = i - 1;
        while (j >= 0 && ev[j] > x) { ev[j + 1] = ev[j]; --j; }
        ev[j + 1] = x;
    }
    // tolerance scale per component
    const double cross = 4.0 * b.hw[0] * b.hw[1];
    const double hmax = 2.0 * fmax(b.hw[0], fmax(b.hw[1], b.hw[2]));
    double scale[4];
    if (mode == 0) { scale[0] = cross; scale[1] = cross * 2.0 * b.hw[0]; scale[2] = cross * 2.0 * b.hw[1]; scale[3] = cross * 2.0 * b.hw[2]; }
    else { scale[0] = hmax; scale[1] = scale[2] = scale[3] = hmax * hmax; }
    const double span = zhi - zlo;
    for (int e = 0; e + 1 < ne; ++e) {
        const double a = ev[e], bb = ev[e + 1];
        if (!(bb - a > 1e-15 * (fabs(a) + fabs(bb) + span))) continue;
        float st0[16], st1[16];   // panels in the smooth-step variable (exact dyadic numbers: float is enough)
        int sp = 1;
        st0[0] = 0.f; st1[0] = 1.f;
        while (sp > 0) {
            --sp;
            const double s0 = st0[sp], s1 = st1[sp];
            const double hs = 0.5 * (s1 - s0), ms = 0.5 * (s1 + s0);
            double K[4] = {0, 0, 0, 0}, G[4] = {0, 0, 0, 0}, v[4];
            for (int j = 0; j < 15; ++j) {
                const int k = j < 8 ? j : 14 - j;
                const double t = j < 8 ? -GK_X(k) : GK_X(k);
                const double s = ms + hs * t;
                const double z = a + (bb - a) * s * s * (3.0 - 2.0 * s);
                const double jac = 6.0 * (bb - a) * s * (1.0 - s) * hs;
                bb_section(b, mode, z, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    K[q] += GK_WK(k) * jac * v[q];
                    if (k & 1) G[q] += GK_WG(k >> 1) * jac * v[q];
                }
            }
            bool ok = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) ok = ok && (fabs(K[q] - G[q]) <= 2e-13 * scale[q] * (s1 - s0));
            if (ok || (s1 - s0) < 1.0 / 4096.0 || sp > 13) {
#pragma unroll
                for (int q = 0; q < 4; ++q) res[q] += K[q];
            } else {
                const float m = 0.5f * (st0[sp] + st1[sp]);
                const float hi = st1[sp];
                st1[sp] = m; ++sp;
                st0[sp] = m; st1[sp] = hi; ++sp;
            }
        }
    }
}

// ---- classification (plain IEEE arithmetic, same order as oracle/geom_oracle.c:classify_in) ----------------------------
// box spans dims mdims[0..m) with bounds lo/hi (positional); optional fixed coordinate (fixd, fixv). 1 full-in, 0 empty, -1 cut
template <int N>
__device__ __forceinline__ int classify_in(const ShapeDev &s, int m, const int *mdims, const double *lo, const double *hi, int fixd, double fixv)
{
    if (s.kind == PB200_LS_HALFSPACE) {
        if (fixd == s.hd) return fixv < s.hc ? 1 : 0;
        for (int q = 0; q < m; ++q)
            if (mdims[q] == s.hd) { if (hi[q] <= s.hc) return 1; if (lo[q] >= s.hc) return 0; return -1; }
        return 0;
    }
    int res = 0;
    for (int b = 0; b < s.nb; ++b) {
        const double *c = s.c + (size_t)b * N;
        const double R2 = __dmul_rn(s.r[b], s.r[b]);
        double dmin2 = 0.0, dmax2 = 0.0;
        for (int q = 0; q < m; ++q) {
            const double cl = c[mdims[q]];
            const double a = __dsub_rn(lo[q], cl), bb = __dsub_rn(cl, hi[q]);
            const double dn = fmax(fmax(a, bb), 0.0);
            const double dx = fmax(fabs(a), fabs(bb));
            dmin2 = __dadd_rn(dmin2, __dmul_rn(dn, dn));
            dmax2 = __dadd_rn(dmax2, __dmul_rn(dx, dx));
        }
        if (fixd >= 0) {
            const double f = __dsub_rn(fixv, c[fixd]);
            dmin2 = __dadd_rn(dmin2, __dmul_rn(f, f));
            dmax2 = __dadd_rn(dmax2, __dmul_rn(f, f));
        }
        if (dmax2 <= R2) return 1;
        if (dmin2 < R2) res = -1;
    }
    return res;
}

__device__ __forceinline__ double node_at(const Grid &g, int d, int j) { return __dadd_rn(g.x0[d], __dmul_rn((double)j + 0.5, g.h[d])); }

// "in"-set measure + first moments about the box centre for boxes of dimension m <= 2 (exact), optional fixed coordinate
template <int N>
__device__ void in_moments_lowdim(const ShapeDev &s, int m, const int *mdims, const double *lo, const double *hi, int fixd, double fixv, double *out)
{
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (s.kind == PB200_LS_HALFSPACE) {
        double l2[3], h2[3], meas = 1.0;
        bool hit = false;
        for (int q = 0; q < m; ++q) { l2[q] = lo[q]; h2[q] = hi[q]; if (mdims[q] == s.hd) { h2[q] = fmin(hi[q], s.hc); hit = true; } }
        if (fixd == s.hd) { if (!(fixv < s.hc)) return; }
        else if (!hit) return;
        for (int q = 0; q < m; ++q) { if (!(h2[q] > l2[q])) return; meas *= h2[q] - l2[q]; }
        out[0] = meas;
        for (int q = 0; q < m; ++q) out[1 + q] = meas * (0.5 * (l2[q] + h2[q]) - 0.5 * (lo[q] + hi[q]));
        return;
    }
    for (int b = 0; b < s.nb; ++b) {
        const double *cb = s.c + (size_t)b * N;
        double R2 = s.r[b] * s.r[b];
        if (fixd >= 0) { const double f = fixv - cb[fixd]; R2 -= f * f; }
        if (m == 0) { if (R2 > 0.0) out[0] += 1.0; continue; }
        if (!(R2 > 0.0)) continue;
        if (m == 1) {
            double len, mom;
            chord_seg(cb[mdims[0]], R2, lo[0], hi[0], 0.5 * (lo[0] + hi[0]), len, mom);
            out[0] += len; out[1] += mom;
        } else {
            const double mx = 0.5 * (lo[0] + hi[0]), my = 0.5 * (lo[1] + hi[1]);
            double o[6];
            disc_rect(cb[mdims[0]] - mx, cb[mdims[1]] - my, sqrt(R2), 0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]), o);
            out[0] += o[0]; out[1] += o[1]; out[2] += o[2];
        }
    }
}

// fluid measure of a box of dimension m <= 2 (any m for half-spaces), sign flip included; type returned; bary optional
template <int N>
__device__ int fluid_lowdim(const ShapeDev &s, int m, const int *mdims, const double *lo, const double *hi, int fixd, double fixv, double &meas,
                            double *bary)
{
    double full = 1.0, mid[3] = {0, 0, 0};
    for (int q = 0; q < m; ++q) { full = __dmul_rn(full, __dsub_rn(hi[q], lo[q])); mid[q] = 0.5 * (lo[q] + hi[q]); }
    int t = classify_in<N>(s, m, mdims, lo, hi, fixd, fixv);
    if (!s.inside && t >= 0) t = 1 - t;
    if (t >= 0) {
        meas = t ? full : 0.0;
        if (bary) for (int q = 0; q < m; ++q) bary[q] = mid[q];
        return t;
    }
    double mom[4];
    in_moments_lowdim<N>(s, m, mdims, lo, hi, fixd, fixv, mom);
    if (!s.inside) { mom[0] = full - mom[0]; mom[1] = -mom[1]; mom[2] = -mom[2]; mom[3] = -mom[3]; }
    meas = mom[0];
    if (bary) for (int q = 0; q < m; ++q) bary[q] = mom[0] > 0.0 ? mid[q] + mom[1 + q] / mom[0] : mid[q];
    return -1;
}

// interface measure + centroid inside an N-box (N <= 2 exact; N == 3 by quadrature in k_geom_cut3d)
template <int N>
__device__ void interface_lowdim(const ShapeDev &s, const double *lo, const double *hi, double &gam, double *cg)
{
    gam = 0.0;
    double mom[3] = {0, 0, 0}, mid[3] = {0, 0, 0};
    for (int q = 0; q < N; ++q) mid[q] = 0.5 * (lo[q] + hi[q]);
    if (s.kind == PB200_LS_HALFSPACE) {
        if (!(lo[s.hd] < s.hc && s.hc < hi[s.hd])) return;
        double meas = 1.0;
        for (int q = 0; q < N; ++q) if (q != s.hd) meas *= hi[q] - lo[q];
        gam = meas;
        for (int q = 0; q < N; ++q) cg[q] = q == s.hd ? s.hc : mid[q];
        return;
    }
    for (int b = 0; b < s.nb; ++b) {
        const double *cb = s.c + (size_t)b * N;
        const double R = s.r[b];
        if (N == 1) {
            const double p0 = cb[0] - R, p1 = cb[0] + R;
            if (p0 >= lo[0] && p0 < hi[0]) { gam += 1.0; mom[0] += p0 - mid[0]; }
            if (p1 >= lo[0] && p1 < hi[0]) { gam += 1.0; mom[0] += p1 - mid[0]; }
        } else if (N == 2) {
            double o[6];
            disc_rect(cb[0] - mid[0], cb[1] - mid[1], R, 0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]), o);
            gam += R * o[3]; mom[0] += R * o[4]; mom[1] += R * o[5];
        }
    }
    if (gam > 0.0) for (int q = 0; q < N; ++q) cg[q] = mid[q] + mom[q] / gam;
}

// ---- kernels ------------------------------------------------------------------------------------------------------------
// local cell ordinal over ALL local planes (ghost planes inside the global domain included): t in [0, nloc)
__device__ __forceinline__ bool local_coords(const Grid &g, int64_t l, int c[PB_MAXD])
{
    // local extents: the slab dimension (the slowest one) has lz planes including the two ghosts, so it must NOT be reduced modulo pd
    int64_t q = l;
    if (g.sd == 0) { c[0] = (int)q; c[1] = 0; c[2] = 0; }
    else if (g.sd == 1) { c[0] = (int)(q % g.pd[0]); c[1] = (int)(q / g.pd[0]); c[2] = 0; }
    else { c[0] = (int)(q % g.pd[0]); q /= g.pd[0]; c[1] = (int)(q % g.pd[1]); c[2] = (int)(q / g.pd[1]); }
    c[g.sd] += g.k0 - 1;
    return c[g.sd] >= 0 && c[g.sd] < g.pd[g.sd];
}

template <int N>
__global__ void k_geom_cells(Grid g, ShapeDev s, int want_cg, GeomOut o, long long *cut_list, int *cut_count, int cut_cap)
{
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < g.nloc; l += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        if (!local_coords(g, l, c)) continue;
        bool real = true;
        for (int d = 0; d < N; ++d) real = real && (c[d] < g.nc[d]);
        if (!real) continue;   // pad layer stays zero (src/capacity.jl:90-92 passes `zero`)
        double lo[PB_MAXD], hi[PB_MAXD];
        const int dims[3] = {0, 1, 2};
        for (int d = 0; d < N; ++d) { lo[d] = node_at(g, d, c[d]); hi[d] = node_at(g, d, c[d] + 1); }
        if (N <= 2 || s.kind == PB200_LS_HALFSPACE) {
            double meas, bary[3];
            const int t = fluid_lowdim<N>(s, N, dims, lo, hi, -1, 0.0, meas, bary);
            o.V[l] = meas; o.ct[l] = (double)t;
            for (int d = 0; d < N; ++d) o.Co[d][l] = bary[d];
            if (t == -1) {
                double gam, cg[3] = {0, 0, 0};
                interface_lowdim<N>(s, lo, hi, gam, cg);
                o.Gam[l] = gam;
                if (want_cg && gam > 0.0) for (int d = 0; d < N; ++d) o.Cg[d][l] = cg[d];
            }
        } else {
            int t = classify_in<N>(s, N, dims, lo, hi, -1, 0.0);
            if (!s.inside && t >= 0) t = 1 - t;
            double full = 1.0;
            for (int d = 0; d < N; ++d) full = __dmul_rn(full, __dsub_rn(hi[d], lo[d]));
            o.ct[l] = (double)t;
            for (int d = 0; d < N; ++d) o.Co[d][l] = 0.5 * (lo[d] + hi[d]);
            if (t >= 0) o.V[l] = t ? full : 0.0;
            else { const int k = atomicAdd(cut_count, 1); if (k < cut_cap) cut_list[k] = l; }
        }
    }
}

__device__ __forceinline__ void make_ballbox(const ShapeDev &s, int b, const double *lo, const double *hi, BallBox &bb)
{
    for (int d = 0; d < 3; ++d) { bb.c[d] = s.c[(size_t)b * 3 + d]; bb.mid[d] = 0.5 * (lo[d] + hi[d]); bb.hw[d] = 0.5 * (hi[d] - lo[d]); }
    bb.R = s.r[b];
}
__device__ __forceinline__ bool ball_touches(const BallBox &bb)
{
    double d2 = 0.0;
    for (int d = 0; d < 3; ++d) { const double a = fmax(fabs(bb.c[d] - bb.mid[d]) - bb.hw[d], 0.0); d2 += a * a; }
    return d2 < bb.R * bb.R;
}

// 3-D cut cells: V, C_omega, Gamma, C_gamma by quadrature of exact sections
__global__ void k_geom_cut3d(Grid g, ShapeDev s, int want_cg, GeomOut o, const long long *cut_list, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t l = cut_list[i];
        int c[PB_MAXD];
        local_coords(g, l, c);
        double lo[3], hi[3], full = 1.0;
        for (int d = 0; d < 3; ++d) { lo[d] = node_at(g, d, c[d]); hi[d] = node_at(g, d, c[d] + 1); full = __dmul_rn(full, __dsub_rn(hi[d], lo[d])); }
        double mom[4] = {0, 0, 0, 0}, sur[4] = {0, 0, 0, 0};
        for (int b = 0; b < s.nb; ++b) {
            BallBox bb;
            make_ballbox(s, b, lo, hi, bb);
            if (!ball_touches(bb)) continue;
            double r[4];
            bb_integrate(bb, 0, r);
            for (int q = 0; q < 4; ++q) mom[q] += r[q];
            bb_integrate(bb, 1, r);
            for (int q = 0; q < 4; ++q) sur[q] += r[q];
        }
        if (!s.inside) { mom[0] = full - mom[0]; mom[1] = -mom[1]; mom[2] = -mom[2]; mom[3] = -mom[3]; }
        o.V[l] = mom[0];
        o.Gam[l] = sur[0];
        for (int d = 0; d < 3; ++d) {
            const double mid = 0.5 * (lo[d] + hi[d]);
            o.Co[d][l] = mom[0] > 0.0 ? mid + mom[1 + d] / mom[0] : mid;
            if (want_cg && sur[0] > 0.0) o.Cg[d][l] = mid + sur[1 + d] / sur[0];
        }
    }
}

// A_d (lower face), B_d (section through the barycentre), W_d (staggered volume between barycentres) -- owned cells
template <int N>
__global__ void k_geom_faces(Grid g, ShapeDev s, GeomOut o, long long *w_list, int *w_count, int w_cap)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < g.nown; t += (int64_t)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const int64_t l = t + g.plane;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            int od[2] = {0, 0}, m = 0;
            bool real_others = true;
            for (int e = 0; e < N; ++e) if (e != d) { od[m++] = e; real_others = real_others && (c[e] < g.nc[e]); }
            if (!real_others) continue;
            double lo[2] = {0, 0}, hi[2] = {0, 0}, meas;
            for (int q = 0; q < m; ++q) { lo[q] = node_at(g, od[q], c[od[q]]); hi[q] = node_at(g, od[q], c[od[q]] + 1); }
            fluid_lowdim<N>(s, m, od, lo, hi, d, node_at(g, d, c[d]), meas, nullptr);
            o.A[d][l] = meas;
            if (c[d] >= g.nc[d]) continue;
            const double ct = o.ct[l];
            double face = 1.0;
            for (int q = 0; q < m; ++q) face = __dmul_rn(face, __dsub_rn(hi[q], lo[q]));
            if (ct == 1.0) o.B[d][l] = face;
            else if (ct == 0.0) o.B[d][l] = 0.0;
            else { fluid_lowdim<N>(s, m, od, lo, hi, d, o.Co[d][l], meas, nullptr); o.B[d][l] = meas; }
            if (c[d] >= 1) {
                double blo[PB_MAXD], bhi[PB_MAXD];
                const int dims[3] = {0, 1, 2};
                for (int e = 0; e < N; ++e) { blo[e] = node_at(g, e, c[e]); bhi[e] = node_at(g, e, c[e] + 1); }
                blo[d] = o.Co[d][l - g.stride[d]];
                bhi[d] = o.Co[d][l];
                if (N <= 2 || s.kind == PB200_LS_HALFSPACE) {
                    fluid_lowdim<N>(s, N, dims, blo, bhi, -1, 0.0, meas, nullptr);
                    o.W[d][l] = meas;
                } else {
                    int tt = classify_in<N>(s, N, dims, blo, bhi, -1, 0.0);
                    if (!s.inside && tt >= 0) tt = 1 - tt;
                    double full = 1.0;
                    for (int e = 0; e < N; ++e) full = __dmul_rn(full, __dsub_rn(bhi[e], blo[e]));
                    if (tt >= 0) o.W[d][l] = tt ? full : 0.0;
                    else { const int k = atomicAdd(w_count, 1); if (k < w_cap) w_list[k] = l * 4 + d; }
                }
            }
        }
    }
}

__global__ void k_geom_w3d(Grid g, ShapeDev s, GeomOut o, const long long *w_list, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t l = w_list[i] >> 2;
        const int d = (int)(w_list[i] & 3);
        int c[PB_MAXD];
        local_coords(g, l, c);
        double lo[3], hi[3], full = 1.0;
        for (int e = 0; e < 3; ++e) { lo[e] = node_at(g, e, c[e]); hi[e] = node_at(g, e, c[e] + 1); }
        lo[d] = o.Co[d][l - g.stride[d]];
        hi[d] = o.Co[d][l];
        for (int e = 0; e < 3; ++e) full = __dmul_rn(full, __dsub_rn(hi[e], lo[e]));
        double vol = 0.0;
        for (int b = 0; b < s.nb; ++b) {
            BallBox bb;
            make_ballbox(s, b, lo, hi, bb);
            if (!ball_touches(bb)) continue;
            double r[4];
            bb_integrate(bb, 0, r);
            vol += r[0];
        }
        o.W[d][l] = s.inside ? vol : full - vol;
    }
}


static int geometry_build(pb200_ctx *ctx, const Grid &g, const pb200_levelset *ls, int compute_centroids, GeomOut &o)
{
    const int N = g.N;
    ShapeDev s;
    s.kind = ls->kind; s.nb = ls->kind == PB200_LS_BALLS ? ls->nballs : 0; s.inside = ls->fluid_inside ? 1 : 0; s.hd = ls->hs_dim; s.hc = ls->hs_c;
    s.c = nullptr; s.r = nullptr;
    double *dc = nullptr, *dr = nullptr;
    if (ls->kind == PB200_LS_BALLS) {
        if (ls->nballs < 1 || !ls->centers || !ls->radii) return set_err(ctx, PB200_EINVAL, "ball level set needs centres and radii");
        CUDA_TRY(ctx, cudaMalloc((void **)&dc, sizeof(double) * (size_t)ls->nballs * N));
        CUDA_TRY(ctx, cudaMalloc((void **)&dr, sizeof(double) * (size_t)ls->nballs));
        CUDA_TRY(ctx, cudaMemcpyAsync(dc, ls->centers, sizeof(double) * (size_t)ls->nballs * N, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(dr, ls->radii, sizeof(double) * (size_t)ls->nballs, cudaMemcpyHostToDevice, ctx->stream));
        s.c = dc; s.r = dr;
    } else if (ls->kind == PB200_LS_HALFSPACE) {
        if (ls->hs_dim < 0 || ls->hs_dim >= N) return set_err(ctx, PB200_EINVAL, "half-space dimension out of range");
    } else return set_err(ctx, PB200_EINVAL, "unknown level-set kind");
    int *d_cnt = nullptr;
    long long *cut_list = nullptr, *w_list = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&d_cnt, 2 * sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), ctx->stream));
    const bool quad3d = N == 3 && ls->kind == PB200_LS_BALLS;
    int cut_cap = 0;
    if (quad3d) {
        // cut cells are an O(n^(2/3)) set; the list is sized generously and overflow is detected and retried
        cut_cap = (int)fmin(2.0e9, fmax(1.0e5, 64.0 * pow((double)g.nloc, 2.0 / 3.0)));
        if ((int64_t)cut_cap > g.nloc) cut_cap = (int)g.nloc;
        CUDA_TRY(ctx, cudaMalloc((void **)&cut_list, sizeof(long long) * (size_t)cut_cap));
    }
    const int grid = red_grid(ctx, g.nloc);
    int h_cnt[2] = {0, 0};
    for (;;) {
        DISPATCH_GEOM_N(N, (k_geom_cells<NN><<<grid, RED_THREADS, 0, ctx->stream>>>(g, s, compute_centroids, o, cut_list, d_cnt, cut_cap)));
        LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_cnt[0] <= cut_cap) break;
        cudaFree(cut_list);
        cut_cap = h_cnt[0];
        CUDA_TRY(ctx, cudaMalloc((void **)&cut_list, sizeof(long long) * (size_t)cut_cap));
        CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), ctx->stream));
    }
    if (quad3d && h_cnt[0] > 0) {
        const int gq = (h_cnt[0] + 63) / 64;
        k_geom_cut3d<<<gq, 64, 0, ctx->stream>>>(g, s, compute_centroids, o, cut_list, h_cnt[0]);
        LAUNCH_CHECK(ctx);
    }
    int w_cap = quad3d ? 8 * h_cnt[0] + 1024 : 0;
    if (quad3d) CUDA_TRY(ctx, cudaMalloc((void **)&w_list, sizeof(long long) * (size_t)w_cap));
    const int gridf = red_grid(ctx, g.nown);
    for (;;) {
        DISPATCH_GEOM_N(N, (k_geom_faces<NN><<<gridf, RED_THREADS, 0, ctx->stream>>>(g, s, o, w_list, d_cnt + 1, w_cap)));
        LAUNCH_CHECK(ctx);
        CUDA_TRY(ctx, cudaMemcpyAsync(h_cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_cnt[1] <= w_cap) break;
        cudaFree(w_list);
        w_cap = h_cnt[1];
        CUDA_TRY(ctx, cudaMalloc((void **)&w_list, sizeof(long long) * (size_t)w_cap));
        CUDA_TRY(ctx, cudaMemsetAsync(d_cnt + 1, 0, sizeof(int), ctx->stream));
    }
    if (quad3d && h_cnt[1] > 0) {
        const int gq = (h_cnt[1] + 63) / 64;
        k_geom_w3d<<<gq, 64, 0, ctx->stream>>>(g, s, o, w_list, h_cnt[1]);
        LAUNCH_CHECK(ctx);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_cnt); cudaFree(cut_list); cudaFree(w_list); cudaFree(dc); cudaFree(dr);
    return PB200_OK;
}
