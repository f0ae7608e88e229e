// geometry.cuh -- cut-cell moments on the device (stub; filled in below)
#pragma once
#include "common.cuh"
struct GeomOut { double *V, *Gam, *ct, *A[PB_MAXD], *B[PB_MAXD], *W[PB_MAXD], *Co[PB_MAXD], *Cg[PB_MAXD]; };
static int geometry_build(pb200_ctx *ctx, const Grid &g, const pb200_levelset *ls, int compute_centroids, GeomOut &o)
{
    return set_err(ctx, PB200_EUNSUPPORTED, "geometry kernels not built yet");
}
