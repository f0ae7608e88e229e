// fold2.cuh -- TMA-staged operator apply of the folded system and the fused CG iteration built on it.
//
// What it replaces: kf_apply_dense (fold.cuh: tile in registers + warp shuffles, neighbours re-loaded per thread) and kf_cg_p.
// Reference rows behind the operator: /root/reference/src/solver/diffusion.jl:212-241, 334-389 (see fold.cuh for the algebra).
//
// Staging.  The Krylov vectors live in a re-pitched copy of the local grid (x pitch = multiple of 32 doubles), which makes them
// describable to the Tensor Memory Accelerator: one `cp.async.bulk.tensor` (SASS: UTMALDG) brings a tile PLUS its one-cell halo --
// a 36 x 34 (2-D) or 36 x 10 x 6 (3-D) box of doubles -- into shared memory and signals an mbarrier; out-of-range coordinates (boxes
// hanging over the array) are zero-filled by the hardware.  A block walks its tile list with TWO stages: the box of tile i+1 is in
// flight while tile i is computed, so the dependent chain record -> address -> data that bounded the register kernel (one round trip
// per tile, 20 us launches at 2048^2) is gone, and every neighbour value is read once from L2 instead of up to 3 times.
//
// Fusion (MODE 5).  The CG search-direction update needs no pass of its own: p_k = z_k + beta_k p_{k-1} is formed IN SHARED MEMORY on the
// tile + halo from the staged boxes of z and p_{k-1} (recomputing the halo ring is 13 % more flops and no extra DRAM traffic), the
// solution update x += alpha_{k-1} p_{k-1} rides along on the tile's own cells, and v = M^ p_k with the partial sum of (p_k, v) follows.
// Per unknown and iteration the CG then moves 48 B here + 24 B in kf_cg_update = 72 B in 2 streaming launches instead of 80 B in 3.
// p is double-buffered (neighbouring blocks still read p_{k-1} halos while this block stores p_k).
//
// Slab-partitioned grids: tiles whose box reaches a ghost plane that a neighbour rank fills ("ghost class") cannot recompute p_k there.
// They take the unfused route on a second stream -- kf2_pupd (pointwise p update) -> halo exchange of p_k -> plain staged apply --
// WHILE the interior class runs the fused kernel: the halo exchange is overlapped with the interior stencil work.
#pragma once
#include <cuda.h>

#include "fold.cuh"

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f2_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void f2_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f2_mbar_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        if (!ok && clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a box that never lands is a bug (bad descriptor), not something to wait out
    } while (!ok);
}
__device__ __forceinline__ void f2_tma_load_2d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((unsigned long long)m), "r"(bar),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void f2_tma_load_3d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"((unsigned long long)m),
                 "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// the same copies with an L2 eviction policy (createpolicy) attached
__device__ __forceinline__ void f2_tma_load_2d_h(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
                 "l"((unsigned long long)m), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void f2_tma_load_3d_h(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
                 "l"((unsigned long long)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ unsigned long long f2_policy_evict_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long f2_policy_evict_first() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void f2_st_hint(double *a, double v, unsigned long long pol) { asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory"); }

template <int N> struct F2Box {
    // x extent 36 = tile 32 + TWO cells on either side: the innermost TMA coordinate must be 16-byte aligned (measured on B200: a box of
    // doubles starting at an odd x raises "illegal instruction", tests/experiments/tma_box_probe.cu), so the box starts at x0 - 2
    static constexpr int HX = 2, BX = 32 + 2 * HX, BY = N == 2 ? 34 : 10, BZ = N == 2 ? 1 : 6;
    static constexpr int NB = BX * BY * BZ;
    static constexpr int BYTES = NB * 8;                           // what one TMA box delivers
    static constexpr int SLOT = (BYTES + 127) / 128 * 128;         // 128-byte aligned shared-memory slot
};
struct alignas(64) F2Maps { CUtensorMap a[2]; CUtensorMap b[2]; };   // per bulk field: a = the staged vector (x, or z in MODE 5), b = p_{k-1} (MODE 5)

// Band heads (one rank).  The interface-band part of an iteration used to be two launches of its own between the streaming kernels
// (kf_apply_band after the apply, kf_band_poly after the residual update): O(band) work, ~13 us of latency each, a third of an iteration at
// 2048^2.  Both are folded into the streaming kernels as a HEAD that every block runs on its share of the band / fringe cells before it starts
// on its tiles -- possible because neither needs anything the same launch produces:
//   * apply head: gathers p_k on a band cell and its neighbours as z + beta p_{k-1} (+ dz on band cells) from the INPUTS of the launch, writes
//     the band couplings of the bulk rows to the compact array ya (the tile part writes the dense part to v: no read-modify-write between blocks),
//     the w rows of v, p_k and x on the interface unknowns (kf2_pupd is gone), and adds its share of (p_k, v) to the block's partial sum;
//   * update head: the residual on the E cells and on the interface unknowns is kept in compact arrays with two buffers by iteration parity, so
//     the head forms r_k = r_{k-1} - alpha (v + ya) on a band cell and its band neighbours from the OLD values while other blocks write the new
//     ones, applies the band polynomial (dz, rho_band) and adds the previous dz to p_k on the band cells (the tile kernel formed p_k without it).
struct BandHead {
    int on, nE, nEp, nB, nbulk;
    const long long *Ecell; const int *EB, *EnbrB, *EnbrE, *EofB, *eord;
    const double *Eblk;
    const long long *Bq; const int *Bidx; const double *Bblk; int nBp;      // band-indexed copies (update head)
    double *ya, *dzw;
    long long ld0, dP, sq[PB_MAXD];
    double ca, cb;          // band polynomial: dz = ca r_B + cb M^_BB r_B
};
__device__ __forceinline__ long long bh_q(const BandHead &b, long long l) { return b.dP ? l + (l / b.ld0) * b.dP : l; }

struct F2Args {
    FVec a, pold, y, pnew, xs, aux;
    BandHead bh;
    const double *dz; const int *bord; int nB;    // band preconditioner correction z_B - r_B (MODE 5, kf2_pupd); nullptr: none
    int sl_old, sl_cur;                           // rho groups of the previous / the current iteration
    StopCrit stop;
    PolyCoef pc;
    double *partials, *results; unsigned *counter;
    const double *res;
    int accumulate;                               // the published dot product is ADDED to the slot (a second launch of the same apply)
    int dbg;                                      // debugging switches (PB200_DBG_F3)
    int l2hint;                                   // L2 eviction hints of the pipelined kernel: 1 boxes evict_last, 2 x tile evict_first, 4 stores evict_first
    const double *off[2][PB_MAXD];                // coefficient arrays of the tiles without constants (reference pitch)
};
// debugging: record an out-of-range index (code = kernel * 100 + site) in res[28] and SKIP the access
#define F2_BAD(A, code) (*(const_cast<double *>((A).res) + 28) = (double)(code))

// MODE 0..4: as kf_apply_dense (0 no dot; 1 (x, y); 2 (aux, y); 3 (y, x), (y, y); 4 y = pc.r aux + pc.z x + pc.A M^ x with (aux, y)).
// MODE 5: fused CG step (header).  `maps` must describe A.a (and A.pold for MODE 5).
template <int N, int MODE>
__global__ void __launch_bounds__(FCH, N == 2 ? 4 : 3) kf2_apply(const __grid_constant__ F2Maps maps, Grid g, FoldDev fd, Items I, F2Args A)
{
    using B = F2Box<N>;
    constexpr int NA = MODE == 5 ? 2 : 1;
    constexpr int STAGE = NA * B::SLOT;
    extern __shared__ __align__(128) unsigned char f2_smem[];
    __shared__ __align__(8) unsigned long long bars[2];
    if (A.stop.sl_rr >= 0 && fold_done(A.res, A.stop)) return;
    const int tid = (int)threadIdx.x, lane = tid & 31, ty = tid >> 5;
    const uint32_t sm0 = f2_smem_u32(f2_smem), bar0 = f2_smem_u32(&bars[0]);
    if (tid == 0) {
        f2_mbar_init(bar0, 1); f2_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double alpha = 0.0, beta = 0.0;
    if (MODE == 5) {
        alpha = A.res[FS_XPEND] != 0.0 ? A.res[FS_ALPHA] : 0.0;                       // x += alpha_{k-1} p_{k-1} (nothing pending in the first iteration)
        beta = safe_div(rho_at(A.res, A.sl_cur), rho_at(A.res, A.sl_old));            // beta_k = rho_k / rho_{k-1}
    }
    auto issue = [&](int it, int stage) {   // one thread: arm the stage's barrier and start the box copies of item `it`
        const TileRec R = I.rec[it];
        const uint32_t dst = sm0 + stage * STAGE, bar = bar0 + 8 * stage;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was read / written through the generic proxy before
        f2_mbar_expect_tx(bar, NA * B::BYTES);
        const CUtensorMap *ma = R.f == 0 ? &maps.a[0] : &maps.a[1];
        if (N == 2) f2_tma_load_2d(dst, ma, bar, R.ox - B::HX, R.oy - 1); else f2_tma_load_3d(dst, ma, bar, R.ox - B::HX, R.oy - 1, R.oz - 1);
        if (NA == 2) {
            const CUtensorMap *mb = R.f == 0 ? &maps.b[0] : &maps.b[1];
            if (N == 2) f2_tma_load_2d(dst + B::SLOT, mb, bar, R.ox - B::HX, R.oy - 1); else f2_tma_load_3d(dst + B::SLOT, mb, bar, R.ox - B::HX, R.oy - 1, R.oz - 1);
        }
    };
    double v[2] = {0.0, 0.0};
    uint32_t phase0 = 0, phase1 = 0;
    int stage = 0;
    if (tid == 0 && (int)blockIdx.x < I.n) issue(blockIdx.x, 0);
    constexpr int TYM = N == 2 ? FU : 1, KY = N == 2 ? 1 : 0, SY = B::BX, SZ = B::BX * B::BY;
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const int nx = it + gridDim.x;
        if (tid == 0 && nx < I.n) issue(nx, stage ^ 1);   // (that stage was released by the __syncthreads that ended the previous round)
        const TileRec R = I.rec[it];
        const unsigned char flags = I.uni[it];
        const bool uni = (flags & 1) != 0;
        const int f = R.f;
        if (stage == 0) { f2_mbar_wait(bar0, phase0); phase0 ^= 1; } else { f2_mbar_wait(bar0 + 8, phase1); phase1 ^= 1; }
        double *__restrict__ sA = reinterpret_cast<double *>(f2_smem + stage * STAGE);
        double *__restrict__ sP = MODE == 5 ? sA + B::SLOT / 8 : sA;   // the vector the stencil runs on
        // this thread's FU cells: box index, global indices, validity
        int bi[FU]; long long l[FU], q[FU]; bool ok[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, l[k], q[k]);
            bi[k] = (lane + B::HX) + (ty * TYM + KY * k + 1) * SY + (N == 3 ? (k + 1) * SZ : 0);
        }
        if (A.dbg & 16) {
#pragma unroll
            for (int k = 0; k < FU; ++k) if (ok[k] && (q[k] < 0 || q[k] >= I.nq)) { F2_BAD(A, 202); ok[k] = false; }
        }
        double pown[FU], xown[FU];
        if (MODE == 5) {
            double *__restrict__ xf = f == 0 ? A.xs.f[0] : A.xs.f[1];
#pragma unroll
            for (int k = 0; k < FU; ++k) { pown[k] = sP[bi[k]]; xown[k] = ok[k] ? xf[q[k]] : 0.0; }
            __syncthreads();   // every thread holds its p_{k-1} values before the box is overwritten
            // p_k = z + beta p_{k-1} on the whole box (tile + halo)
            for (int j = tid; j < B::NB; j += FCH) sP[j] = sA[j] + beta * sP[j];
            if (A.dz != nullptr && (flags & 2)) {   // tiles that hold band cells: z = r + dz there
                __syncthreads();
                for (int j = tid; j < B::NB; j += FCH) {
                    const int jx = j % B::BX, jy = (j / B::BX) % B::BY, jz = j / (B::BX * B::BY);
                    const long long gx = R.ox + jx - B::HX, gy = R.oy + jy - 1, gz = N == 3 ? R.oz + jz - 1 : 0;
                    if (gx >= 0 && gx < I.ld0 && gy >= 0 && gy < I.ld1 && gz >= 0 && gz < I.ld2) {
                        const int bo = A.bord[gx + I.ld0 * (gy + I.ld1 * gz)];
                        if (bo >= 0) sP[j] += A.dz[(size_t)f * A.nB + bo];
                    }
                }
            }
            __syncthreads();
        }
        const double *__restrict__ uc = I.ucoef + (size_t)it * PB_MAXD;
        const double *__restrict__ of0 = f == 0 ? fd.off[0][0] : fd.off[1][0];
        const double *__restrict__ of1 = f == 0 ? fd.off[0][1] : fd.off[1][1];
        const double *__restrict__ of2 = f == 0 ? fd.off[0][N > 2 ? 2 : 0] : fd.off[1][N > 2 ? 2 : 0];
        double *__restrict__ yf = f == 0 ? A.y.f[0] : A.y.f[1];
        const double *__restrict__ af = f == 0 ? A.aux.f[0] : A.aux.f[1];
        double c[FU], acc[FU], av[FU];
        if (uni) {
            const double cx = uc[0], cy = uc[1], cz = N == 3 ? uc[2] : 0.0;
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                c[k] = sP[bi[k]];
                acc[k] = c[k] + cx * (sP[bi[k] - 1] + sP[bi[k] + 1]) + cy * (sP[bi[k] - SY] + sP[bi[k] + SY]);
                if (N == 3) acc[k] += cz * (sP[bi[k] - SZ] + sP[bi[k] + SZ]);
            }
        } else {
            double cm[FU][N], cp[FU][N];
#pragma unroll
            for (int k = 0; k < FU; ++k) {
#pragma unroll
                for (int d = 0; d < N; ++d) {
                    const double *__restrict__ of = d == 0 ? of0 : (d == 1 ? of1 : of2);
                    if ((A.dbg & 16) && ok[k] && (l[k] < 0 || l[k] + g.stride[d] >= I.nl)) { F2_BAD(A, 201); ok[k] = false; }
                    cm[k][d] = ok[k] ? of[l[k]] : 0.0;
                    cp[k][d] = ok[k] ? of[l[k] + g.stride[d]] : 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                c[k] = sP[bi[k]];
                acc[k] = c[k] + cm[k][0] * sP[bi[k] - 1] + cp[k][0] * sP[bi[k] + 1] + cm[k][1] * sP[bi[k] - SY] + cp[k][1] * sP[bi[k] + SY];
                if (N == 3) acc[k] += cm[k][N - 1] * sP[bi[k] - SZ] + cp[k][N - 1] * sP[bi[k] + SZ];
            }
        }
        if (MODE == 2 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < FU; ++k) av[k] = ok[k] ? af[q[k]] : 0.0;
        }
        if (MODE == 5) {
            double *__restrict__ pn = f == 0 ? A.pnew.f[0] : A.pnew.f[1];
            double *__restrict__ xf = f == 0 ? A.xs.f[0] : A.xs.f[1];
#pragma unroll
            for (int k = 0; k < FU; ++k)
                if (ok[k]) {
                    yf[q[k]] = acc[k];
                    pn[q[k]] = c[k];
                    xf[q[k]] = xown[k] + alpha * pown[k];
                    v[0] += c[k] * acc[k];
                }
        } else {
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                if (MODE == 4) acc[k] = A.pc.r * av[k] + A.pc.z * c[k] + A.pc.A * acc[k];
                if (ok[k]) {
                    yf[q[k]] = acc[k];
                    if (MODE == 1) v[0] += c[k] * acc[k];
                    if (MODE == 2 || MODE == 4) v[0] += av[k] * acc[k];
                    if (MODE == 3) { v[0] += acc[k] * c[k]; v[1] += acc[k] * acc[k]; }
                }
            }
        }
        __syncthreads();   // the stage may be refilled
        stage ^= 1;
    }
    if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5) { double w[1] = {v[0]}; block_reduce_publish<1>(w, A.partials, A.results, A.counter, A.accumulate != 0); }
    if (MODE == 3) block_reduce_publish<2>(v, A.partials, A.results, A.counter, A.accumulate != 0);
}

// Pointwise search-direction + solution update on a list of items: the ghost-class tiles of a slab-partitioned grid (whose p_k must exist
// before the halo exchange) and the compact interface unknowns w (which the staged kernel does not touch).
//   p_k = z + beta p_{k-1} (z = a + dz on band cells),  x += alpha_{k-1} p_{k-1}
__global__ void __launch_bounds__(FCH) kf2_pupd(Items I, F2Args A)
{
    if (A.stop.sl_rr >= 0 && fold_done(A.res, A.stop)) return;
    const double alpha = A.res[FS_XPEND] != 0.0 ? A.res[FS_ALPHA] : 0.0;
    const double beta = safe_div(rho_at(A.res, A.sl_cur), rho_at(A.res, A.sl_old));
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        const int f = R.f;
        const double *__restrict__ zf = f == 0 ? A.a.f[0] : (f == 1 ? A.a.f[1] : A.a.f[2]);
        const double *__restrict__ po = f == 0 ? A.pold.f[0] : (f == 1 ? A.pold.f[1] : A.pold.f[2]);
        double *__restrict__ pn = f == 0 ? A.pnew.f[0] : (f == 1 ? A.pnew.f[1] : A.pnew.f[2]);
        double *__restrict__ xf = f == 0 ? A.xs.f[0] : (f == 1 ? A.xs.f[1] : A.xs.f[2]);
        const bool band_tile = A.dz != nullptr && (f == 2 || (I.uni[it] & 2));
        long long i[FU], iq[FU]; bool ok[FU]; double pv[FU], zv[FU], xv[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, i[k], iq[k]);
            if (ok[k]) { pv[k] = po[iq[k]]; xv[k] = xf[iq[k]]; zv[k] = zf[iq[k]]; }
        }
        if (band_tile) {
#pragma unroll
            for (int k = 0; k < FU; ++k)
                if (ok[k]) {
                    const long long bo = f == 2 ? i[k] : (long long)A.bord[i[k]];
                    if (bo >= 0) zv[k] += A.dz[(size_t)f * A.nB + bo];
                }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k)
            if (ok[k]) {
                pn[iq[k]] = zv[k] + beta * pv[k];
                xf[iq[k]] = xv[k] + alpha * pv[k];
            }
    }
}

// CG residual update of the fused iteration: r -= alpha_k v, publishes (rho, rr); records alpha_k and the pending solution update for the
// NEXT fused apply (or kf2_xflush).  Same arithmetic as kf_cg_update.
// carry: single rank -- a skipped iteration copies the (rho, rr, rho_band, rho_poly_band) group forward here (several ranks: kf2_carry after the allreduce)
__global__ void __launch_bounds__(FCH) kf2_update(Items I, double *res, int sl_rho, int sl_new, int carry, FVec qv, FVec r, double *partials, unsigned *counter, StopCrit stop)
{
    if (fold_done(res, stop)) {
        if (carry && blockIdx.x == 0 && threadIdx.x == 0) { res[sl_new] = res[sl_rho]; res[sl_new + 1] = res[sl_rho + 1]; res[sl_new + 2] = res[sl_rho + 2]; res[sl_new + 3] = res[sl_rho + 3]; }
        return;
    }
    const double alpha = safe_div(rho_at(res, sl_rho), res[FS_SIG_D] + res[FS_SIG_B] + res[FS_SIG_G]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { res[FS_ITERS] += 1.0; res[FS_ALPHA] = alpha; res[FS_XPEND] = 1.0; }
    double v[2] = {0.0, 0.0};
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        const int f = R.f;
        const double *__restrict__ qf = f == 0 ? qv.f[0] : (f == 1 ? qv.f[1] : qv.f[2]);
        double *__restrict__ rf = f == 0 ? r.f[0] : (f == 1 ? r.f[1] : r.f[2]);
        long long i[FU], iq[FU]; bool ok[FU]; double vv[FU], rv[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, i[k], iq[k]);
            if (ok[k]) { vv[k] = qf[iq[k]]; rv[k] = rf[iq[k]]; }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k)
            if (ok[k]) {
                const double rn = rv[k] - alpha * vv[k];
                rf[iq[k]] = rn;
                v[0] += rn * rn;
            }
    }
    v[1] = v[0];
    block_reduce_publish<2>(v, partials, res + sl_new, counter);
}
// compact copy of the residual on the E cells (both parity buffers), once per solve
__global__ void kf2_gather_rE(BandHead b, FVec r, double *rE0, double *rE1)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < b.nE; e += gridDim.x * blockDim.x) {
        const long long lq = bh_q(b, b.Ecell[e]);
        for (int f = 0; f < b.nbulk; ++f) { const double v = r.f[f][lq]; rE0[(size_t)f * b.nEp + e] = v; rE1[(size_t)f * b.nEp + e] = v; }
    }
}
// Residual update of the fused iteration WITH the update head (see BandHead).  The dense residual is updated IN PLACE (a second dense buffer
// pushed the iteration's working set at 2048^2 from 170 to 204 MB against 126 MB of L2: 23 -> 32 us); what the head reads of r_{k-1} while other
// blocks overwrite it -- the bulk entries of the E cells and the interface unknowns -- lives in compact arrays with two buffers by iteration
// parity (rEold / rEnew, rold.f[2] / r.f[2]).  v = qv + ya on the E cells; publishes (rho, rr, rho_band).
// HB > 0: the first HB blocks of the grid run the head (one thread per band cell) and nothing else, the others stream the tiles meanwhile -- with
// every block running its share of the head first, nothing streamed during the head's dependent gathers (2048^2: 23 -> 33 us).  HB = 0: every block does both.
template <int N, int LPC>
__global__ void __launch_bounds__(FCH, 4) kf2_update_b(Items I, double *res, int sl_rho, int sl_new, FVec qv, FVec rold, FVec r, const double *__restrict__ rEold,
                                                       double *__restrict__ rEnew, FVec pnew, BandHead b, int HB, int carry, double *partials, unsigned *counter, StopCrit stop)
{
    if (fold_done(res, stop)) {
        if (carry && blockIdx.x == 0 && threadIdx.x == 0) { res[sl_new] = res[sl_rho]; res[sl_new + 1] = res[sl_rho + 1]; res[sl_new + 2] = res[sl_rho + 2]; res[sl_new + 3] = res[sl_rho + 3]; }
        return;
    }
    const double alpha = safe_div(rho_at(res, sl_rho), res[FS_SIG_D] + res[FS_SIG_B] + res[FS_SIG_G]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { res[FS_ITERS] += 1.0; res[FS_ALPHA] = alpha; res[FS_XPEND] = 1.0; }
    double v[3] = {0.0, 0.0, 0.0};
    const size_t ES = (size_t)b.nEp;
    const bool two = b.nbulk > 1;
    const bool head_block = HB == 0 || (int)blockIdx.x < HB;
    if (b.nB > 0 && head_block) {
        // ---- head: band polynomial on r_k, formed from r_{k-1}, v and ya; LPC lanes per band cell (lane `sub` takes the blocks k = sub, sub + LPC, ...) ----
        const int nhb = HB == 0 ? (int)gridDim.x : HB;
        const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), sub = gtid % LPC, gid = gtid / LPC, ngroups = nhb * (int)blockDim.x / LPC;
        const int rounds = (b.nB + ngroups - 1) / ngroups;
        for (int rd = 0; rd < rounds; ++rd) {
            const int bo_raw = rd * ngroups + gid;
            const bool live = bo_raw < b.nB;
            const int bo = live ? bo_raw : b.nB - 1;        // (groups past the end redo the last cell: the shuffles are warp-wide)
            const size_t BS = (size_t)b.nBp;
            const int e_raw2 = b.Bidx[bo];
            const bool ghost_row = e_raw2 < 0;            // (band cell of a ghost plane: no row here; the lanes still run for the warp-wide shuffles)
            const int e = ghost_row ? 0 : e_raw2;
            const long long lq = b.Bq[bo];
            const double dz0 = b.dzw[bo], dz1 = two ? b.dzw[(size_t)b.nB + bo] : 0.0;      // (lane 0 needs them after the reduction: in flight with the rest)
            double r0 = 0.0, r1 = 0.0, r2 = 0.0, x0 = 0.0, x1 = 0.0, xw = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < 1 + 2 * N; k0 += LPC) {
                const int k = k0 + sub;
                if (LPC > 1 && k >= 1 + 2 * N) break;
                long long ln = lq;
                int nb = bo, en = e;
                if (k > 0) {
                    const int kk = k - 1, d = kk >> 1;
                    ln = (kk & 1) ? lq + b.sq[d] : lq - b.sq[d];
                    nb = ghost_row ? -1 : b.Bidx[(size_t)(1 + kk) * BS + bo];
                    en = b.Bidx[(size_t)(1 + 2 * N + kk) * BS + bo];
                }
                // branch-free: a neighbour that is no band cell reads this cell's entries and counts with weight 0 (all loads of a lane are in flight together)
                const double wgt = nb >= 0 ? 1.0 : 0.0;
                if (nb < 0) { nb = bo; en = e; ln = lq; }
                const double *__restrict__ c = b.Bblk + (size_t)(k * 9) * BS + bo;
                const double c0 = c[0], c1 = c[BS], c2 = c[2 * BS], c3 = c[3 * BS], c4 = c[4 * BS], c5 = c[5 * BS], c6 = c[6 * BS], c7 = c[7 * BS], c8 = c[8 * BS];
                const double v0 = wgt * fma(-alpha, qv.f[0][ln] + b.ya[en], rEold[en]);
                const double v1 = two ? wgt * fma(-alpha, qv.f[1][ln] + b.ya[ES + en], rEold[ES + en]) : 0.0;
                const double v2 = wgt * fma(-alpha, qv.f[2][nb], rold.f[2][nb]);
                if (k == 0) { x0 = v0; x1 = v1; xw = v2; }
                r0 += c0 * v0 + c1 * v1 + c2 * v2;
                r1 += c3 * v0 + c4 * v1 + c5 * v2;
                r2 += c6 * v0 + c7 * v1 + c8 * v2;
            }
            if (LPC > 1) {
#pragma unroll
                for (int o = LPC / 2; o > 0; o >>= 1) {
                    r0 += __shfl_xor_sync(0xffffffffu, r0, o, LPC); r1 += __shfl_xor_sync(0xffffffffu, r1, o, LPC); r2 += __shfl_xor_sync(0xffffffffu, r2, o, LPC);
                }
            }
            if (sub != 0 || !live || ghost_row) continue;                 // (lane 0 of the group took k = 0: it holds x0, x1, xw)
            const double a0 = r0 + x0, a1 = r1 + x1, a2 = r2 + xw;        // (I + band block) r_B
            // p_k on the band cell lacks the band correction the tile kernel could not see: add the one it was formed with, then replace it
            pnew.f[0][lq] += dz0;
            if (two) pnew.f[1][lq] += dz1;
            const double o0 = b.ca * x0 + b.cb * a0, o1 = b.ca * x1 + b.cb * a1, o2 = b.ca * xw + b.cb * a2;
            b.dzw[bo] = o0;
            b.dzw[(size_t)b.nB + bo] = o1;
            b.dzw[2 * (size_t)b.nB + bo] = o2;
            v[2] += x0 * o0 + x1 * o1 + xw * o2;
        }
    }
    const int tb = HB == 0 ? (int)blockIdx.x : (int)blockIdx.x - HB, nt = HB == 0 ? (int)gridDim.x : (int)gridDim.x - HB;
    for (int it = tb; it >= 0 && it < I.n; it += nt) {
        const TileRec R = I.rec[it];
        const int f = R.f;
        const bool hasE = f < 2 && (I.uni[it] & 8) != 0;
        const double *__restrict__ qf = f == 0 ? qv.f[0] : (f == 1 ? qv.f[1] : qv.f[2]);
        const double *__restrict__ ro = f == 0 ? rold.f[0] : (f == 1 ? rold.f[1] : rold.f[2]);
        double *__restrict__ rf = f == 0 ? r.f[0] : (f == 1 ? r.f[1] : r.f[2]);
        long long i[FU], iq[FU]; bool ok[FU]; double vv[FU], rv[FU]; int ee[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, i[k], iq[k]);
            ee[k] = -1;
            if (ok[k]) { vv[k] = qf[iq[k]]; rv[k] = ro[iq[k]]; }
        }
        if (hasE) {
#pragma unroll
            for (int k = 0; k < FU; ++k)
                if (ok[k]) {
                    ee[k] = b.eord[i[k]];
                    if (ee[k] >= 0) vv[k] += b.ya[(size_t)f * ES + ee[k]];
                }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k)
            if (ok[k]) {
                const double rn = fma(-alpha, vv[k], rv[k]);
                rf[iq[k]] = rn;
                if (hasE && ee[k] >= 0) rEnew[(size_t)f * ES + ee[k]] = rn;      // the compact copy the NEXT update head reads (two buffers by parity)
                v[0] += rn * rn;
            }
    }
    v[1] = v[0];
    block_reduce_publish<3>(v, partials, res + sl_new, counter);
}
// skipped iteration: carry the whole (rho, rr, rho_band, rho_poly_band) group forward (the fused iteration has no kf_cg_p to do it)
__global__ void kf2_carry(double *res, int sl_old, int sl_new, StopCrit stop)
{
    if (threadIdx.x == 0 && fold_done(res, stop)) {
        res[sl_new] = res[sl_old]; res[sl_new + 1] = res[sl_old + 1]; res[sl_new + 2] = res[sl_old + 2]; res[sl_new + 3] = res[sl_old + 3];
    }
}

// after the last iteration: the solution still lacks alpha_k p_k (the fused apply of iteration k+1 would have added it)
__global__ void __launch_bounds__(FCH) kf2_xflush(Items I, const double *res, FVec p0, FVec p1, FVec xs)
{
    if (res[FS_XPEND] == 0.0) return;
    const double alpha = res[FS_ALPHA];
    const bool odd = (((long long)(res[FS_ITERS] + 0.5)) & 1) != 0;   // iteration j reads P[j & 1] and writes P[(j + 1) & 1]
    FV_LOOP(I) {
        (void)i;
        const double *__restrict__ pf = odd ? p1.f[f] : p0.f[f];
        xs.f[f][q] += alpha * pf[q];
    }
}

// =================================================================================================================================
// kf3_apply -- the interior constant-coefficient tiles (every cell valid, one constant per direction: 80-88 % of the cells of the
// benchmark problems), as a warp-specialised TMA pipeline.
//
// Measured with kf2_apply above (one box in flight per block, operands of the epilogue loaded with ordinary loads): ~3 us per tile and
// block whatever the box size -- the chain  tile record -> wait for the box -> x / coefficient loads -> stores  is a series of exposed
// memory round trips, and four resident blocks do not hide it (0.28 of the HBM peak at 2048^2, 0.26 at 512^3).  Here
//   * ONE producer thread (warp 8) runs S stages ahead: it reads the tile record, copies what the consumers need of it (store offset,
//     field, the stencil constants) into the stage header, arms the stage's FULL barrier and issues every read of the tile as a TMA
//     copy -- the box of the staged vector, the box of p_{k-1} (MODE 5), the 32 x 32 / 32 x 8 x 4 tile of x (MODE 5) or of aux;
//   * EIGHT consumer warps wait on FULL, compute from shared memory only, store v / p_k / x straight from registers (256-byte rows)
//     and release the stage through its EMPTY barrier (one arrival per warp).  No block-wide barrier in the loop: p_k = z + beta p_{k-1}
//     is formed on the fly for the cell and its neighbours (a column of FU + 2 values per thread is shared along the k direction).
// General tiles (interface band, border ring, partial tiles) keep kf2_apply; they run beside this kernel on the second stream.
// =================================================================================================================================
// apply head (see BandHead): groups of LPC lanes walk the E list; `gtid` / `nthr` = index / number of the threads of the whole grid that run heads
template <int N, int LPC>
__device__ __forceinline__ void f3_band_head(const F2Args &A, double alpha, double beta, double &sig, int gtid, int nthr)
{
    const BandHead &b = A.bh;
    if (b.nE <= 0) return;
    const size_t ES = (size_t)b.nEp;
    const bool two = b.nbulk > 1;
    const int sub = gtid % LPC, gid = gtid / LPC, ngroups = nthr / LPC;
    const int rounds = (b.nE + ngroups - 1) / ngroups;
    const double *__restrict__ z0 = A.a.f[0], *__restrict__ z1 = A.a.f[1], *__restrict__ zw = A.a.f[2];
    const double *__restrict__ q0 = A.pold.f[0], *__restrict__ q1 = A.pold.f[1], *__restrict__ qw = A.pold.f[2];
    const double *__restrict__ dz = b.dzw;
    for (int rd = 0; rd < rounds; ++rd) {
        const int e_raw = rd * ngroups + gid;
        const bool live = e_raw < b.nE;
        const int e = live ? e_raw : b.nE - 1;      // (groups past the end redo the last cell: the shuffles are warp-wide)
        const long long lq = bh_q(b, b.Ecell[e]);
        const int bo = b.EB[e];
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, x0 = 0.0, x1 = 0.0, xw = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 1 + 2 * N; k0 += LPC) {
            const int k = k0 + sub;
            if (LPC > 1 && k >= 1 + 2 * N) break;
            long long ln = lq;
            int nb = bo;
            if (k > 0) {
                const int kk = k - 1, d = kk >> 1;
                ln = (kk & 1) ? lq + b.sq[d] : lq - b.sq[d];
                nb = b.EnbrB[(size_t)kk * ES + e];
            }
            const bool nonzero = k == 0 ? bo >= 0 : (bo >= 0 || nb >= 0);     // (kf_blocks: other blocks are exactly zero)
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (nonzero || k == 0) {
                v0 = fma(beta, q0[ln], z0[ln]);
                if (two) v1 = fma(beta, q1[ln], z1[ln]);
                if (nb >= 0) {
                    const double e0 = dz[nb], e1 = two ? dz[(size_t)b.nB + nb] : 0.0;
                    v0 += e0; v1 += e1;
                    v2 = fma(beta, qw[nb], zw[nb] + dz[2 * (size_t)b.nB + nb]);
                    if (k == 0) { d0 = e0; d1 = e1; }
                }
            }
            if (k == 0) { x0 = v0; x1 = v1; xw = v2; }
            if (nonzero) {
                const double *__restrict__ c = b.Eblk + (size_t)(k * 9) * ES + e;
                r0 += c[0] * v0 + c[ES] * v1 + c[2 * ES] * v2;
                r1 += c[3 * ES] * v0 + c[4 * ES] * v1 + c[5 * ES] * v2;
                r2 += c[6 * ES] * v0 + c[7 * ES] * v1 + c[8 * ES] * v2;
            }
        }
        if (LPC > 1) {
#pragma unroll
            for (int o = LPC / 2; o > 0; o >>= 1) {
                r0 += __shfl_xor_sync(0xffffffffu, r0, o, LPC); r1 += __shfl_xor_sync(0xffffffffu, r1, o, LPC); r2 += __shfl_xor_sync(0xffffffffu, r2, o, LPC);
            }
        }
        if (sub != 0 || !live) continue;       // (lane 0 of the group took k = 0: it holds x0, x1, xw, d0, d1)
        // bulk rows: the tile kernel writes v = p' (band cell: its dense couplings are zero, p' = p_k - dz) or the dense stencil (fringe cell)
        b.ya[e] = d0 + r0;
        if (two) b.ya[ES + e] = d1 + r1;
        sig += d0 * (x0 - d0) + d1 * (x1 - d1) + x0 * (d0 + r0) + x1 * (d1 + r1);
        if (bo >= 0) {
            const double yw = xw + r2;
            A.y.f[2][bo] = yw;
            A.pnew.f[2][bo] = xw;
            A.xs.f[2][bo] += alpha * qw[bo];
            sig += xw * yw;
        }
    }
}

// Which list item a block works on in its n-th round.  run == 1: item n * grid + block (neighbouring items on neighbouring blocks at the same time).
// run > 1 (3-D lists, sorted so that consecutive items are z / y neighbours): the list is cut into runs of `run` items, run j goes to block j mod grid,
// so a block walks spatial neighbours back to back and the halo planes its next box shares with the last one are still in L2.  (ncu, 1024 x 1024 x 128:
// with item-wise striding the boxes were fetched from DRAM 2.1 times over -- blocks drift apart, and a neighbour's box comes tens of microseconds later.)
__device__ __forceinline__ int f3_item(int n, int run, int grid, int block) { return run <= 1 ? n * grid + block : ((n / run) * grid + block) * run + n % run; }

struct alignas(64) F3Maps { CUtensorMap a[2], b[2], t[2]; };   // boxes of the staged vector and of p_{k-1}; tile (no halo) of x (MODE 5) / aux (MODE 2, 4)
struct F3Hdr {
    long long baseq, base;            // tile origin in the re-pitched Krylov vectors / in the reference-pitch arrays (coefficients, band map)
    double cx, cy, cz;                // the tile's constants (flags & 1)
    int f, flags;                     // field; bit 0 constant coefficients, bit 1 holds band cells, bit 2 every cell valid
    int ox, oy, oz;
    short nx; signed char ylo, yhi, zlo, zhi;
};
template <int N> struct F3Tile { static constexpr int BYTES = FTILE * 8; };

__device__ __forceinline__ void f3_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// has_t: the third map is valid (MODE 5 always; MODE 2 / 4: aux is a vector of its own -- otherwise aux == the staged vector)
// BH: 0 no band head; 1, 2: apply head with that many lanes per band / fringe cell (MODE 5, one rank)
template <int N, int MODE, int S, int BH = 0>
__global__ void __launch_bounds__(FCH + 32, 2) kf3_apply(const __grid_constant__ F3Maps maps, Items I, F2Args A, int has_t, int dbg)
{
    using B = F2Box<N>;
    constexpr int NBX = MODE == 5 ? 2 : 1;
    constexpr bool TT = MODE == 5 || MODE == 2 || MODE == 4;               // a tile slot exists in the stage
    constexpr int STAGE = NBX * B::SLOT + (TT ? F3Tile<N>::BYTES : 0);
    extern __shared__ unsigned char f3_raw[];
    __shared__ __align__(8) unsigned long long full[S], empty[S];
    __shared__ F3Hdr hdr[S];
    if (A.stop.sl_rr >= 0 && fold_done(A.res, A.stop)) return;
    const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t raw = f2_smem_u32(f3_raw), sm0 = (raw + 127u) & ~127u;   // TMA destinations must be 128-byte aligned
    unsigned char *smem = f3_raw + (sm0 - raw);
    const uint32_t fb = f2_smem_u32(&full[0]), eb = f2_smem_u32(&empty[0]);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { f2_mbar_init(fb + 8 * s, 1); f2_mbar_init(eb + 8 * s, FCH / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double v[2] = {0.0, 0.0};
    if (wid == FCH / 32) {
        // ---- producer ----------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            const bool use_t = TT && (MODE == 5 || has_t);
            const uint32_t bytes = NBX * B::BYTES + (use_t ? F3Tile<N>::BYTES : 0);
            const unsigned long long pol_last = f2_policy_evict_last(), pol_first = f2_policy_evict_first();
            int n = 0;
            TileRec Rn;
            const int G = (int)gridDim.x, bk = (int)blockIdx.x;
            if (f3_item(0, I.run, G, bk) < I.n) Rn = I.rec[f3_item(0, I.run, G, bk)];
            for (int it = f3_item(0, I.run, G, bk); it < I.n; it = f3_item(++n, I.run, G, bk)) {
                const TileRec R = Rn;
                const double c0 = I.ucoef[(size_t)it * PB_MAXD], c1 = I.ucoef[(size_t)it * PB_MAXD + 1], c2 = I.ucoef[(size_t)it * PB_MAXD + 2];
                { const int nx = f3_item(n + 1, I.run, G, bk); if (nx < I.n) Rn = I.rec[nx]; }      // next record: in flight while this tile is issued
                const int s = n % S, k = n / S;
                if (k > 0) f2_mbar_wait(eb + 8 * s, (uint32_t)((k - 1) & 1));      // the consumers have released the stage's previous tile
                F3Hdr h;
                h.baseq = R.baseq; h.base = R.base; h.cx = c0; h.cy = c1; h.cz = c2; h.f = R.f;
                h.flags = (int)I.uni[it] | (R.full ? 4 : 0);
                h.ox = R.ox; h.oy = R.oy; h.oz = R.oz; h.nx = R.nx; h.ylo = R.ylo; h.yhi = R.yhi; h.zlo = R.zlo; h.zhi = R.zhi;
                hdr[s] = h;
                const uint32_t dst = sm0 + s * STAGE, bar = fb + 8 * s;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                f2_mbar_expect_tx(bar, bytes);                                       // (release: the header is visible to whoever passes the barrier)
                const int fi = R.f == 0 ? 0 : 1;
                if (A.l2hint & 1) {
                    if (N == 2) f2_tma_load_2d_h(dst, &maps.a[fi], bar, R.ox - B::HX, R.oy - 1, pol_last); else f2_tma_load_3d_h(dst, &maps.a[fi], bar, R.ox - B::HX, R.oy - 1, R.oz - 1, pol_last);
                    if (MODE == 5) {
                        if (N == 2) f2_tma_load_2d_h(dst + B::SLOT, &maps.b[fi], bar, R.ox - B::HX, R.oy - 1, pol_last);
                        else f2_tma_load_3d_h(dst + B::SLOT, &maps.b[fi], bar, R.ox - B::HX, R.oy - 1, R.oz - 1, pol_last);
                    }
                } else {
                    if (N == 2) f2_tma_load_2d(dst, &maps.a[fi], bar, R.ox - B::HX, R.oy - 1); else f2_tma_load_3d(dst, &maps.a[fi], bar, R.ox - B::HX, R.oy - 1, R.oz - 1);
                    if (MODE == 5) {
                        if (N == 2) f2_tma_load_2d(dst + B::SLOT, &maps.b[fi], bar, R.ox - B::HX, R.oy - 1);
                        else f2_tma_load_3d(dst + B::SLOT, &maps.b[fi], bar, R.ox - B::HX, R.oy - 1, R.oz - 1);
                    }
                }
                if (use_t) {
                    if (A.l2hint & 2) { if (N == 2) f2_tma_load_2d_h(dst + NBX * B::SLOT, &maps.t[fi], bar, R.ox, R.oy, pol_first); else f2_tma_load_3d_h(dst + NBX * B::SLOT, &maps.t[fi], bar, R.ox, R.oy, R.oz, pol_first); }
                    else { if (N == 2) f2_tma_load_2d(dst + NBX * B::SLOT, &maps.t[fi], bar, R.ox, R.oy); else f2_tma_load_3d(dst + NBX * B::SLOT, &maps.t[fi], bar, R.ox, R.oy, R.oz); }
                }
            }
        }
    } else {
        // ---- consumers ---------------------------------------------------------------------------------------------------------
        double alpha = 0.0, beta = 0.0;
        if (MODE == 5) {
            alpha = A.res[FS_XPEND] != 0.0 ? A.res[FS_ALPHA] : 0.0;
            beta = safe_div(rho_at(A.res, A.sl_cur), rho_at(A.res, A.sl_old));
        }
        if (BH > 0 && MODE == 5) f3_band_head<N, BH>(A, alpha, beta, v[0], (int)blockIdx.x * FCH + tid, (int)gridDim.x * FCH);   // (the producer is already prefetching tiles)
        constexpr int TYM = N == 2 ? FU : 1, SY = B::BX, SZ = B::BX * B::BY;
        constexpr int SK = N == 2 ? SY : SZ;                  // box stride of the k direction (the FU cells of a thread)
        const int ty = wid;
        const bool use_t = TT && (MODE == 5 || has_t);
        const unsigned long long pol_st = f2_policy_evict_first();
        int n = 0;
        for (int it = f3_item(0, I.run, (int)gridDim.x, (int)blockIdx.x); it < I.n; it = f3_item(++n, I.run, (int)gridDim.x, (int)blockIdx.x)) {
            const int s = n % S, k0 = n / S;
            f2_mbar_wait(fb + 8 * s, (uint32_t)(k0 & 1));
            const F3Hdr h = hdr[s];
            if (dbg & 16) {   // debugging: the header must be the one of this tile, and the staged boxes must hold this tile's data
                const TileRec Rc = I.rec[it];
                if (h.baseq != Rc.baseq || h.f != Rc.f) atomicAdd(const_cast<double *>(A.res) + 30, 1.0);
                const double *__restrict__ sAc = reinterpret_cast<const double *>(smem + s * STAGE);
                const int bb = (lane + B::HX) + (ty * (N == 2 ? FU : 1) + 1) * B::BX + (N == 3 ? B::BX * B::BY : 0);
                const double *__restrict__ ga = Rc.f == 0 ? A.a.f[0] : A.a.f[1];
                const long long qq = Rc.baseq + lane + (long long)(ty * (N == 2 ? FU : 1)) * I.P0;
                if (sAc[bb] != ga[qq]) atomicAdd(const_cast<double *>(A.res) + 31, 1.0);
            }
            const double *__restrict__ sA = reinterpret_cast<const double *>(smem + s * STAGE);
            const double *__restrict__ sB = sA + B::SLOT / 8;
            const double *__restrict__ sT = sA + NBX * (B::SLOT / 8);
            const int b0 = (lane + B::HX) + (ty * TYM + 1) * SY + (N == 3 ? SZ : 0);   // box index of the thread's cell k = 0
            const int t0 = lane + (ty * TYM) * 32;                                    // tile index of it (2-D: rows; 3-D: plane k adds 256)
            constexpr int TK = N == 2 ? 32 : 256;
            if ((h.flags & 7) != 5) {
                // ---- general tile: streamed coefficients and / or partial validity and / or band cells -----------------------------------
                const int f = h.f;
                // (band cells: p = r + beta p_old is formed WITHOUT the band correction dz here; kf_apply_band adds it -- their dense couplings are zero)
                const bool uni = (h.flags & 1) != 0;
                const double *__restrict__ of0 = f == 0 ? A.off[0][0] : A.off[1][0];
                const double *__restrict__ of1 = f == 0 ? A.off[0][1] : A.off[1][1];
                const double *__restrict__ of2 = f == 0 ? A.off[0][N > 2 ? 2 : 0] : A.off[1][N > 2 ? 2 : 0];
                const long long l0 = h.base + lane + (long long)(ty * TYM) * I.ld0, q0g = h.baseq + lane + (long long)(ty * TYM) * I.P0;
                const long long st1 = I.ld0, st2 = I.ld0 * I.ld1;
                bool ok[FU];
                double cm[FU][N], cp[FU][N];
#pragma unroll
                for (int k = 0; k < FU; ++k) {
                    const int yr = ty * TYM + (N == 2 ? k : 0), zr = N == 3 ? k : 0;
                    ok[k] = lane < h.nx && yr >= h.ylo && yr < h.yhi && zr >= h.zlo && zr < h.zhi;
                    const long long l = l0 + (long long)k * I.ustride;
#pragma unroll
                    for (int d = 0; d < N; ++d) {
                        const double *__restrict__ of = d == 0 ? of0 : (d == 1 ? of1 : of2);
                        const long long sd = d == 0 ? 1 : (d == 1 ? st1 : st2);
                        const double cu = d == 0 ? h.cx : (d == 1 ? h.cy : h.cz);
                        cm[k][d] = !ok[k] ? 0.0 : (uni ? cu : of[l]);
                        cp[k][d] = !ok[k] ? 0.0 : (uni ? cu : of[l + sd]);
                    }
                }
#define F3_AT(j) (MODE == 5 ? sA[(j)] + beta * sB[(j)] : sA[(j)])
                double *__restrict__ yf = f == 0 ? A.y.f[0] : A.y.f[1];
#pragma unroll
                for (int k = 0; k < FU; ++k) {
                    const int b = b0 + k * SK;
                    const double c = F3_AT(b);
                    double a = c + cm[k][0] * F3_AT(b - 1) + cp[k][0] * F3_AT(b + 1) + cm[k][1] * F3_AT(b - SY) + cp[k][1] * F3_AT(b + SY);
                    if (N == 3) a += cm[k][N - 1] * F3_AT(b - SZ) + cp[k][N - 1] * F3_AT(b + SZ);
                    if (ok[k]) {
                        const long long q = q0g + (long long)k * I.ustrideq;
                        if (MODE == 5) {
                            double *__restrict__ pn = f == 0 ? A.pnew.f[0] : A.pnew.f[1];
                            double *__restrict__ xf = f == 0 ? A.xs.f[0] : A.xs.f[1];
                            yf[q] = a;
                            pn[q] = c;
                            xf[q] = sT[t0 + k * TK] + alpha * sB[b];
                            v[0] += c * a;
                        } else {
                            const double av = (MODE == 2 || MODE == 4) ? (use_t ? sT[t0 + k * TK] : c) : 0.0;
                            if (MODE == 4) a = A.pc.r * av + A.pc.z * c + A.pc.A * a;
                            yf[q] = a;
                            if (MODE == 1) v[0] += c * a;
                            if (MODE == 2 || MODE == 4) v[0] += av * a;
                            if (MODE == 3) { v[0] += a * c; v[1] += a * a; }
                        }
                    }
                }
#undef F3_AT
                __syncwarp();
                if (lane == 0) f3_mbar_arrive(eb + 8 * s);
                continue;
            }
            // the staged vector on the thread's column (k = -1 .. FU) and on the neighbours of its FU cells
#define F3_AT(j) (MODE == 5 ? sA[(j)] + beta * sB[(j)] : sA[(j)])
            double col[FU + 2];
#pragma unroll
            for (int k = -1; k <= FU; ++k) col[k + 1] = F3_AT(b0 + k * SK);
            double acc[FU];
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                const int b = b0 + k * SK;
                double a = col[k + 1] + h.cx * (F3_AT(b - 1) + F3_AT(b + 1));
                if (N == 2) a += h.cy * (col[k] + col[k + 2]);
                else a += h.cy * (F3_AT(b - SY) + F3_AT(b + SY)) + h.cz * (col[k] + col[k + 2]);
                acc[k] = a;
            }
#undef F3_AT
            const long long q0 = h.baseq + lane + (long long)(ty * TYM) * I.P0;
            const int f = h.f;
            double *__restrict__ yf = f == 0 ? A.y.f[0] : A.y.f[1];
            if ((dbg & 16) && MODE == 5 && N == 3) {   // debugging: recompute the thread's four cells from GLOBAL memory
                const double *__restrict__ gz = f == 0 ? A.a.f[0] : A.a.f[1];
                const double *__restrict__ gp = f == 0 ? A.pold.f[0] : A.pold.f[1];
                const double *__restrict__ uc = I.ucoef + (size_t)it * PB_MAXD;
                for (int k = 0; k < FU; ++k) {
                    const long long q = q0 + (long long)k * I.ustrideq, sy = I.P0, sz = I.ustrideq;
#define GP(j) (gz[(j)] + beta * gp[(j)])
                    const double ref = GP(q) + uc[0] * (GP(q - 1) + GP(q + 1)) + uc[1] * (GP(q - sy) + GP(q + sy)) + uc[2] * (GP(q - sz) + GP(q + sz));
#undef GP
                    if (fabs(ref - acc[k]) > 1e-12 * (1.0 + fabs(ref))) atomicAdd(const_cast<double *>(A.res) + 29, 1.0);
                }
            }
            if ((dbg & 16) && (q0 < 0 || q0 + (long long)(FU - 1) * I.ustrideq >= I.nq)) { F2_BAD(A, 301); __syncwarp(); if (lane == 0) f3_mbar_arrive(eb + 8 * s); continue; }
            if (MODE == 5) {
                double *__restrict__ pn = f == 0 ? A.pnew.f[0] : A.pnew.f[1];
                double *__restrict__ xf = f == 0 ? A.xs.f[0] : A.xs.f[1];
#pragma unroll
                for (int k = 0; k < FU; ++k) {
                    const long long q = q0 + (long long)k * I.ustrideq;
                    const double pold = sB[b0 + k * SK], xo = (dbg & 1) ? xf[q] : sT[t0 + k * TK];
                    if (A.l2hint & 4) { f2_st_hint(yf + q, acc[k], pol_st); f2_st_hint(pn + q, col[k + 1], pol_st); f2_st_hint(xf + q, xo + alpha * pold, pol_st); }
                    else {
                        yf[q] = acc[k];
                        pn[q] = col[k + 1];
                        xf[q] = xo + alpha * pold;
                    }
                    v[0] += col[k + 1] * acc[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < FU; ++k) {
                    const long long q = q0 + (long long)k * I.ustrideq;
                    const double av = (MODE == 2 || MODE == 4) ? (use_t ? sT[t0 + k * TK] : col[k + 1]) : 0.0;
                    double a = acc[k];
                    if (MODE == 4) a = A.pc.r * av + A.pc.z * col[k + 1] + A.pc.A * a;
                    yf[q] = a;
                    if (MODE == 1) v[0] += col[k + 1] * a;
                    if (MODE == 2 || MODE == 4) v[0] += av * a;
                    if (MODE == 3) { v[0] += a * col[k + 1]; v[1] += a * a; }
                }
            }
            if (dbg & 2) asm volatile("bar.sync 1, 256;" ::: "memory");
            __syncwarp();
            if (lane == 0) f3_mbar_arrive(eb + 8 * s);    // this warp has read everything it needs of the stage
        }
    }
    if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5) { double w[1] = {v[0]}; block_reduce_publish<1>(w, A.partials, A.results, A.counter, A.accumulate != 0); }
    if (MODE == 3) block_reduce_publish<2>(v, A.partials, A.results, A.counter, A.accumulate != 0);
}

// debugging aid (PB200_DBG_CHECK): compare two vectors on the tiles of a list; out[0] = number of differing cells, out[1] = item index of the last one seen,
// out[2] = its flags (uni | full << 4 | ghost << 5), out[3] = max |a - b|
__global__ void __launch_bounds__(FCH) kf2_dbg_compare(Items I, FVec a, FVec b, double tol, double *out)
{
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        if (R.f >= 2) continue;
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            long long i, q;
            if (!tile_cell(I, R, k, i, q)) continue;
            const double d = fabs(a.f[R.f][q] - b.f[R.f][q]);
            if (d > tol) {
                atomicAdd(out, 1.0);
                out[1] = (double)it; out[2] = (double)(I.uni[it] | (R.full << 4) | (R.ghost << 5));
                if (d > out[3]) out[3] = d;
                atomicMin((int *)(out + 4), it);
                atomicMax((int *)(out + 5), it);
                if (k == 0 && threadIdx.x == 0) atomicAdd(out + 6, 1.0);   // tiles (counted at their first cell)
            }
        }
    }
}
