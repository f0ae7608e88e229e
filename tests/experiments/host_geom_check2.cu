#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <vector>
#define PB_HOST_GEOM 1
#include "../../penguin.jl_b200/csrc/geometry.cuh"
extern "C" int pgo_capacity(int N, const int *ncell, const double *x0, const double *L, int kind, int nb, const double *centers, const double *radii, int inside, int hd, double hc,
                 double *V, double *Gamma, double *ctype, double *A, double *B, double *W, double *Com, double *Cga);
int main()
{
    int nc[2] = {33, 47}; double x0[2] = {0, 0}, L[2] = {4.0, 3.0}, c[2] = {2.01, 1.53}, r = 0.93;
    int pd0 = 34, pd1 = 48; size_t n = (size_t)pd0 * pd1;
    std::vector<double> V(n), G(n), ct(n), A(2 * n), B(2 * n), W(2 * n), Co(2 * n), Cg(2 * n);
    pgo_capacity(2, nc, x0, L, 0, 1, c, &r, 1, 0, 0.0, V.data(), G.data(), ct.data(), A.data(), B.data(), W.data(), Co.data(), Cg.data());
    double hx = L[0] / nc[0], hy = L[1] / nc[1];
    for (int j = 0; j < nc[1]; ++j) for (int i = 0; i < nc[0]; ++i) {
        double lo[2] = {x0[0] + (i + 0.5) * hx, x0[1] + (j + 0.5) * hy}, hi[2] = {x0[0] + (i + 1.5) * hx, x0[1] + (j + 1.5) * hy};
        double mx = 0.5 * (lo[0] + hi[0]), my = 0.5 * (lo[1] + hi[1]);
        double o[6];
        disc_rect(c[0] - mx, c[1] - my, r, 0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]), o);
        size_t idx = i + (size_t)pd0 * j;
        double gam = ct[idx] == -1.0 ? r * o[3] : 0.0;
        if (fabs(gam - G[idx]) > 1e-12 || (ct[idx] == -1.0 && fabs(o[0] - V[idx]) > 1e-13))
            printf("cell (%d,%d) ct %g: gam dev %.17g oracle %.17g | V dev %.17g oracle %.17g | X0=%.17g Y0=%.17g hx=%.17g hy=%.17g\n", i, j, ct[idx], gam, G[idx], o[0], V[idx], c[0] - mx,
                   c[1] - my, 0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]));
    }
    return 0;
}
