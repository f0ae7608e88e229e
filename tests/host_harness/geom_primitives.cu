// Host harness (test infrastructure): compiles the geometry primitives of penguin.jl_b200/csrc/geometry.cuh -- the code the CUDA capacity
// kernels run -- for the CPU (they are __host__ __device__) and compares them with the C oracle (oracle/geom_oracle.c) on random
// configurations.  Output: one line per disagreement (full precision, machine readable) and a summary; tests/test_device_geometry_on_host.py
// builds it with nvcc (host code only runs; no GPU needed), parses the output and arbitrates the disagreements with mpmath.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#define PB_HOST_GEOM 1
#include "../../penguin.jl_b200/csrc/geometry.cuh"
extern "C" void pgo_ball_box(int m, const double *c, double R, const double *lo, const double *hi, double *out);
extern "C" void pgo_sphere_box(int m, const double *c, double R, const double *lo, const double *hi, double *out);
static double urand() { return rand() / (double)RAND_MAX; }

static int run2d(int n)
{
    int bad = 0;
    for (int it = 0; it < n; ++it) {
        const double hx = 0.02 + 0.3 * urand(), hy = 0.02 + 0.3 * urand();
        const double R = it % 3 == 0 ? 0.05 + 0.1 * urand() : 0.3 + 3.0 * urand();
        const double ang = 6.28318 * urand();
        double dist = R + (urand() - 0.5) * 2.2 * hypot(hx, hy);
        if (it % 7 == 0) dist = urand() * R;
        const double c[2] = {dist * cos(ang), dist * sin(ang)};
        const double lo[2] = {-hx, -hy}, hi[2] = {hx, hy};
        double o[6], vb[3], sb[3];
        disc_rect(c[0], c[1], R, hx, hy, o);
        pgo_ball_box(2, c, R, lo, hi, vb);
        pgo_sphere_box(2, c, R, lo, hi, sb);
        const double sc = 4 * hx * hy, h = fmax(hx, hy);
        const double e0 = fabs(o[0] - vb[0]) / sc, e1 = fabs(o[1] - vb[1]) / (sc * h), e2 = fabs(o[2] - vb[2]) / (sc * h);
        const double g0 = fabs(R * o[3] - sb[0]) / h, g1 = fabs(R * o[4] - sb[1]) / (h * h), g2 = fabs(R * o[5] - sb[2]) / (h * h);
        if (e0 > 1e-12 || e1 > 1e-12 || e2 > 1e-12 || g0 > 1e-11 || g1 > 1e-11 || g2 > 1e-11) {
            if (bad < 64)
                printf("BAD2D %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", it, c[0], c[1], R, hx, hy, o[0], vb[0], R * o[3], sb[0]);
            ++bad;
        }
    }
    printf("SUMMARY2D %d %d\n", n, bad);
    return bad;
}

static int run3d(int n)
{
    int bad = 0;
    double worst[2] = {0.0, 0.0};
    for (int it = 0; it < n; ++it) {
        BallBox b;
        for (int q = 0; q < 3; ++q) { b.hw[q] = 0.02 + 0.2 * urand(); b.mid[q] = 0.0; }
        b.R = it % 3 == 0 ? 0.08 + 0.1 * urand() : 0.4 + 2.0 * urand();
        const double diag = sqrt(b.hw[0] * b.hw[0] + b.hw[1] * b.hw[1] + b.hw[2] * b.hw[2]);
        double dist = b.R + (urand() - 0.5) * 2.2 * diag;
        if (it % 7 == 0) dist = urand() * b.R;
        const double th = acos(2.0 * urand() - 1.0), ph = 6.28318 * urand();
        b.c[0] = dist * sin(th) * cos(ph); b.c[1] = dist * sin(th) * sin(ph); b.c[2] = dist * cos(th);
        double lo[3], hi[3];
        for (int q = 0; q < 3; ++q) { lo[q] = -b.hw[q]; hi[q] = b.hw[q]; }
        double v[4], s[4], vo[4], so[4];
        bb_integrate(b, 0, v);
        bb_integrate(b, 1, s);
        pgo_ball_box(3, b.c, b.R, lo, hi, vo);
        pgo_sphere_box(3, b.c, b.R, lo, hi, so);
        const double vol = 8 * b.hw[0] * b.hw[1] * b.hw[2], h = 2 * fmax(b.hw[0], fmax(b.hw[1], b.hw[2]));
        double ev = fabs(v[0] - vo[0]) / vol, es = fabs(s[0] - so[0]) / (h * h);
        for (int q = 1; q < 4; ++q) { ev = fmax(ev, fabs(v[q] - vo[q]) / (vol * h)); es = fmax(es, fabs(s[q] - so[q]) / (h * h * h)); }
        worst[0] = fmax(worst[0], ev); worst[1] = fmax(worst[1], es);
        if (ev > 1e-11 || es > 1e-10) {
            if (bad < 32) printf("BAD3D %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g V %.17g %.17g S %.17g %.17g\n", it, b.c[0], b.c[1], b.c[2], b.R, b.hw[0], b.hw[1], b.hw[2], v[0], vo[0], s[0], so[0]);
            ++bad;
        }
    }
    printf("SUMMARY3D %d %d %.3e %.3e\n", n, bad, worst[0], worst[1]);
    // closed forms: a ball inside the box, and a ball centred on a face / an edge / a corner of a large box
    int badc = 0;
    const double R = 0.37;
    for (int k = 0; k < 4; ++k) {
        BallBox b;
        b.R = R;
        for (int q = 0; q < 3; ++q) { b.hw[q] = 1.0; b.mid[q] = 0.0; b.c[q] = (q < k) ? 1.0 : 0.1 * (q + 1); }   // k coordinates on the box boundary
        double v[4], s[4];
        bb_integrate(b, 0, v);
        bb_integrate(b, 1, s);
        const double frac = 1.0 / (1 << k);
        const double Vex = frac * 4.0 / 3.0 * M_PI * R * R * R, Sex = frac * 4.0 * M_PI * R * R;
        printf("CLOSED %d %.17g %.17g %.17g %.17g\n", k, v[0], Vex, s[0], Sex);
        if (fabs(v[0] - Vex) > 1e-13 || fabs(s[0] - Sex) > 1e-12) ++badc;
    }
    return bad + badc;
}

// the reference's darcy_test.jl geometry: circle centre (0.5, 0.5) r 0.5 on an h = 0.1 grid -- tangent to x = 0 at the grid node (0, 0.5)
static void run_tangent()
{
    const double h = 0.1;
    for (int j = 3; j <= 6; ++j)
        for (int i = 0; i <= 1; ++i) {
            const double lox = i * h, hix = (i + 1) * h, loy = j * h, hiy = (j + 1) * h;
            const double mx = 0.5 * (lox + hix), my = 0.5 * (loy + hiy);
            double o[6];
            disc_rect(0.5 - mx, 0.5 - my, 0.5, 0.5 * (hix - lox), 0.5 * (hiy - loy), o);
            printf("TANGENT %d %d %.17g %.17g %.17g %.17g %.17g %.17g\n", i, j, lox, hix, loy, hiy, o[0], 0.5 * o[3]);
        }
}

int main(int argc, char **argv)
{
    srand(1);
    const int n2 = argc > 1 ? atoi(argv[1]) : 200000, n3 = argc > 2 ? atoi(argv[2]) : 4000;
    run2d(n2);
    run3d(n3);
    run_tangent();
    return 0;
}
