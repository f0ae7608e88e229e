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
    int nc[2] = {40, 40}; double x0[2] = {0, 0}, L[2] = {4.0, 4.0};
    double cen[6] = {1.0, 1.0, 3.0, 1.2, 2.0, 3.0}, rad[3] = {0.6, 0.45, 0.7};
    int pd0 = 41, pd1 = 41; size_t n = (size_t)pd0 * pd1;
    std::vector<double> V(n), G(n), ct(n), A(2 * n), B(2 * n), W(2 * n), Co(2 * n), Cg(2 * n);
    pgo_capacity(2, nc, x0, L, 0, 3, cen, rad, 1, 0, 0.0, V.data(), G.data(), ct.data(), A.data(), B.data(), W.data(), Co.data(), Cg.data());
    double hx = L[0] / nc[0], hy = L[1] / nc[1];
    double sumG = 0, sumV = 0;
    for (int j = 0; j < nc[1]; ++j) for (int i = 0; i < nc[0]; ++i) {
        double lo[2] = {x0[0] + (i + 0.5) * hx, x0[1] + (j + 0.5) * hy}, hi[2] = {x0[0] + (i + 1.5) * hx, x0[1] + (j + 1.5) * hy};
        double mx = 0.5 * (lo[0] + hi[0]), my = 0.5 * (lo[1] + hi[1]);
        size_t idx = i + (size_t)pd0 * j;
        double gam = 0, vol = 0;
        for (int b = 0; b < 3; ++b) {
            double o[6];
            disc_rect(cen[2 * b] - mx, cen[2 * b + 1] - my, rad[b], 0.5 * (hi[0] - lo[0]), 0.5 * (hi[1] - lo[1]), o);
            gam += rad[b] * o[3]; vol += o[0];
        }
        if (ct[idx] != -1.0) gam = 0;
        sumG += G[idx]; sumV += V[idx];
        if (fabs(gam - G[idx]) > 1e-12 || (ct[idx] == -1.0 && fabs(vol - V[idx]) > 1e-13))
            printf("cell (%d,%d) ct %g: gam dev %.17g oracle %.17g | V dev %.17g oracle %.17g\n", i, j, ct[idx], gam, G[idx], vol, V[idx]);
    }
    printf("oracle sum Gamma %.15g analytic %.15g ; sum V %.15g analytic %.15g\n", sumG, 2 * M_PI * (0.6 + 0.45 + 0.7), sumV, M_PI * (0.36 + 0.2025 + 0.49));
    return 0;
}
