// host harness: runs the device geometry primitives of geometry.cuh on the CPU (compiled with -DPB_HOST_GEOM) against the oracle
#include <cstdio>
#include <cmath>
#include <cstdlib>
#define PB_HOST_GEOM 1
#include "../../penguin.jl_b200/csrc/geometry.cuh"
extern "C" void pgo_ball_box(int m, const double *c, double R, const double *lo, const double *hi, double *out);
extern "C" void pgo_sphere_box(int m, const double *c, double R, const double *lo, const double *hi, double *out);
static double urand() { return rand() / (double)RAND_MAX; }
int main()
{
    srand(1);
    int bad = 0;
    for (int it = 0; it < 200000; ++it) {
        double hx = 0.02 + 0.3 * urand(), hy = 0.02 + 0.3 * urand();
        double R = it % 3 == 0 ? 0.05 + 0.1 * urand() : 0.3 + 3.0 * urand();
        double ang = 6.28318 * urand(), dist = R + (urand() - 0.5) * 2.2 * hypot(hx, hy);
        if (it % 7 == 0) dist = urand() * R;
        double c[2] = {dist * cos(ang), dist * sin(ang)};
        double lo[2] = {-hx, -hy}, hi[2] = {hx, hy};
        double o[6], vb[3], sb[3];
        disc_rect(c[0], c[1], R, hx, hy, o);
        pgo_ball_box(2, c, R, lo, hi, vb);
        pgo_sphere_box(2, c, R, lo, hi, sb);
        double sc = 4 * hx * hy, h = fmax(hx, hy);
        double e0 = fabs(o[0] - vb[0]) / sc, e1 = fabs(o[1] - vb[1]) / (sc * h), e2 = fabs(o[2] - vb[2]) / (sc * h);
        double g0 = fabs(R * o[3] - sb[0]) / h, g1 = fabs(R * o[4] - sb[1]) / (h * h), g2 = fabs(R * o[5] - sb[2]) / (h * h);
        if (e0 > 1e-12 || e1 > 1e-12 || e2 > 1e-12 || g0 > 1e-11 || g1 > 1e-11 || g2 > 1e-11) {
            if (bad < 15) printf("it %d c=(%.17g,%.17g) R=%.17g hx=%.17g hy=%.17g | area %g vs %g | arc %g vs %g | errs %g %g %g %g %g %g\n", it, c[0], c[1], R, hx, hy, o[0], vb[0], R * o[3], sb[0], e0, e1, e2, g0, g1, g2);
            ++bad;
        }
    }
    printf("bad = %d\n", bad);
    return 0;
}
