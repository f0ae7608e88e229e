// fold.cuh -- the fast path of the linear solve: the reduced system of /root/reference/src/solver/diffusion.jl:30-43,
// 104-144, 212-241, 334-389 (after remove_zero_rows_cols!, src/solver.jl:59-78, and after the elimination of the known
// unknowns) FOLDED into a symmetric, block-Jacobi-scaled stencil with unit diagonal, and the Krylov kernels that run on it.
//
// Algebra (DESIGN.md "Folded system").  Per phase p the reference operator [G H]' W! [G H] is a sum over faces of rank-1
// terms W!_f g_f g_f' on the four unknowns touching face f = (d, i):
//     g_f = ( a = e_i B_i  on u_i,   b = -B_{i-1}  on u_{i-1},   c = e_i (A_i - B_i)  on gamma_i,   d = -(A_i - B_{i-1})  on gamma_{i-1} )
// (operators.cuh derives the same numbers row-wise).  With the interface unknown w (mono: T_gamma; diph: T_gamma2 and
// T_gamma1 = (g + a2 w) / a1, i.e. gamma_1 = kappa w + known, kappa = a2/a1) and the positive row scalings
//     bulk rows of phase p:  s_p / (c D_p),   s = 1 (mono),  s_1 = b1 / kappa, s_2 = b2 (diph);   interface row: 1 (diph), 1 / (c2 beta) (mono)
// the system matrix is  M = diag(s_p cV V_p / (c D_p)) (+ (alpha/beta) Gamma on w, mono Robin)  +  sum_p s_p sum_f W!_f g'_f g'_f^T,
// symmetric positive definite.  Cut cells couple u_1, u_2 and w of the same cell strongly (small-cell stiffness), so M is scaled
// by the Cholesky factors L_i of its per-cell 3x3 diagonal blocks:  M^ = L^-1 M L^-T  has identity diagonal blocks; outside the
// interface band L_i is diagonal and M^ is a plain (2N+1)-point stencil with unit diagonal and N coefficient arrays per phase.
// Everything that involves a band cell (w active) is kept in compact per-cell 3x3 blocks over the band and its face neighbours.
//
// Data layout: bulk fields are the dense padded arrays of common.cuh (only ACTIVE 256-cell chunks are ever touched: the two
// phases of a diphasic problem are complementary, so streaming both dense would double the traffic); w lives in a compact
// array over the band cells, sorted by cell index (ghost-plane entries form a prefix / suffix, so its halo is two contiguous ranges).
#pragma once
#include <cuda.h>
#include <map>

#include <thrust/execution_policy.h>
#include <thrust/sort.h>

#include "assemble.cuh"
#include "krylov.cuh"
#include "operators.cuh"

#define FCH 256   // cells per chunk == threads per block

struct FoldDev {
    int N, nbulk, has_w;
    double s[2], kap[2];
    double mwc;    // (alpha / beta) Gamma on the w diagonal (mono Robin), else 0
    double wrow;   // row factor of the interface row
    double cVc;    // cV / c
    double c;      // theta dt (1 steady)
    const unsigned char *m[2];   // bulk masks per phase
    const unsigned char *mw;     // mask carrying MB_IFREE
    PhaseDev ph[2];
    double *sc[2];
    double *off[2][PB_MAXD];
    int nB, nBlo, nBown, nE;
    const long long *Bcell, *Ecell;
    const int *bord;
    double *Linv;    // [5][nB]: i00 i11 i20 i21 i22
    int *EB;         // [nE]
    int *EnbrB;      // [2N][nEp]        (SoA: one THREAD works on one band / fringe cell, consecutive threads on consecutive cells)
    double *Eblk;    // [(1+2N)*9][nEp]  block 0: self, 1+2d: lower neighbour in d, 2+2d: upper; each 3x3 row-major
    unsigned char *Efix;   // [nE] bit k: gathered cell k (0 self, 1 + kk neighbours) lies in a tile of the fused kernel (its p lacks the band correction dz)
    // band work folded into the two streaming kernels of the fused CG iteration (fold2.cuh "band heads", one rank):
    const int *eord;       // [nloc] index of a cell in the E list (band + fringe cells), -1 elsewhere
    const int *EofB;       // [nB]   E index of band cell k
    const int *EnbrE;      // [2N][nEp] E index of the face neighbours that are band cells (-1: not a band cell), beside EnbrB
    // the same rows indexed by BAND cell (update head: one dependent load less per gather)
    const long long *Bq;   // [nB] re-pitched index of band cell k
    const int *Bidx;       // [(1 + 4N)][nBp]: row 0 E index of the cell, rows 1 + kk band index of neighbour kk, rows 1 + 2N + kk its E index
    const double *Bblk;    // [(1+2N)*9][nBp] copy of the cell's Eblk rows
    int nBp;
    double *ya;            // [2][nEp] band-coupling part of v = M^ p on the bulk rows of the E cells (the tile kernel writes the dense part to v)
    int nEp;         // nE rounded up to a multiple of 32
    // Krylov vectors (FVec bulk fields) live in a RE-PITCHED copy of the local grid: x rows padded from ld0 (the reference's odd n+1) to P0,
    // a multiple of 32 doubles, so that every 32-cell tile row is 256-byte aligned and the arrays can be described to TMA (global
    // strides must be multiples of 16 bytes).  They never cross the ABI; capacities, masks, states keep the reference pitch.
    int fix_lo, fix_hi;          // local planes of the slab dimension whose tiles the fused kernel processes (kf_blocks -> Efix)
    long long ld0, dP;           // reference pitch, P0 - ld0
    long long sq[PB_MAXD];       // strides of the re-pitched layout (1, P0, P0 * ld1)
};
// re-pitched index of local cell l
__device__ __forceinline__ long long fd_q(const FoldDev &fd, long long l) { return fd.dP ? l + (l / fd.ld0) * fd.dP : l; }

struct FVec { double *f[3]; };   // f[0], f[1]: dense bulk fields; f[2]: compact w

// Work items.  A bulk item is one TILE of 1024 cells of one bulk field, processed by a 256-thread block, FU = 4 cells per thread:
// 1024 x 1 (1-D: thread t takes x = t + 256 k), 32 x 32 (2-D: row ty + 8 k), 32 x 8 x 4 (3-D: plane k) in the local array (ghost planes
// included).  A warp reads 256 contiguous bytes per row, a thread has four independent cells in flight, and only the outermost ring of
// tiles touches the domain border.  A w item is 1024 consecutive entries of the compact interface array.  Only tiles that hold a free
// unknown are listed; each carries a precomputed record so that no index arithmetic beyond shifts happens in the kernels.
#define FU 4
#ifndef PB_APPLY_GP
#define PB_APPLY_GP(N) 2   // cells of a general tile in flight per thread in kf_apply_dense (4: more registers, no gain measured)
#endif
#define FTILE (FCH * FU)
struct __align__(16) TileRec {
    long long base;        // local linear index of the tile origin (w items: first entry)
    long long baseq;       // the same cell in the re-pitched layout of the Krylov vectors (w items: = base)
    int ox, oy, oz;        // tile origin in local array coordinates (TMA box coordinates)
    short nx;              // valid extent along x (w items: valid entries)
    signed char f;         // field: 0, 1 bulk, 2 w
    signed char ylo, yhi;  // valid range of the tile-relative y coordinate
    signed char zlo, zhi;  // valid range of the tile-relative z coordinate
    signed char full;      // 1: every cell of the tile is valid (interior tile)
    signed char ghost;     // 1: the tile + halo box touches a ghost plane whose data comes from a neighbour rank
    signed char pad_;
};
struct Items {
    const int *it; int n;
    int run;                       // kf3_apply: items are handed to the blocks in runs of this many consecutive items (0 / 1: item by item), see f3_item
    const TileRec *rec;            // [n]
    int shx;                       // log2 of the thread extent in x: 8 (1-D, w items use this layout too) or 5
    int kx, ky, kz;                // tile-relative coordinate advance per k: (256,0,0) 1-D, (0,1,0) 2-D, (0,0,1) 3-D
    int tym;                       // thread row ty covers tile rows ty * tym + ky * k  (2-D: FU consecutive rows per thread; else 1)
    long long ustride;             // linear index advance per k
    long long P0, ustrideq;        // x pitch and per-k advance of the re-pitched Krylov vectors
    long long nq, nl;              // sizes of the re-pitched / reference-pitch arrays (debug bounds checks)
    long long ld0, ld1, ld2;       // local array extents
    int T0, T1, T2, nt0, nt1;      // tile extents and tiles per direction (build-time only)
    int sd, lz;                    // slab dimension and its local extent (first / last plane of it are ghosts)
    int glo, ghi;                  // a neighbour rank exists below / above (its data fills that ghost plane)
    int wlo, whi;
    const unsigned char *uni;      // [n] bit 0: the tile's coefficients are constants (ucoef); bit 1: the tile holds band cells; bit 3: holds E cells (kf_tile_meta)
    const double *ucoef;           // [n][PB_MAXD]
};
// cell k (0..FU-1) of this thread inside tile R: linear index (reference pitch), index in the re-pitched Krylov vectors, validity
__device__ __forceinline__ bool tile_cell(const Items &I, const TileRec &R, int k, long long &idx, long long &q)
{
    if (R.f >= 2) {   // compact w: 1-D layout
        const int xr = (int)threadIdx.x + FCH * k;
        idx = R.base + xr;
        q = idx;
        return xr < R.nx;
    }
    const int tx = (int)threadIdx.x & ((1 << I.shx) - 1), ty = (int)threadIdx.x >> I.shx;
    const int xr = tx + I.kx * k, yr = ty * I.tym + I.ky * k, zr = I.kz * k;
    idx = R.base + tx + (long long)(ty * I.tym) * I.ld0 + (long long)k * I.ustride;
    q = R.baseq + tx + (long long)(ty * I.tym) * I.P0 + (long long)k * I.ustrideq;
    return xr < R.nx && yr >= R.ylo && yr < R.yhi && zr >= R.zlo && zr < R.zhi;
}
__device__ __forceinline__ bool tile_cell(const Items &I, const TileRec &R, int k, long long &idx)
{
    long long q;
    return tile_cell(I, R, k, idx, q);
}
__device__ __forceinline__ long long tile_of_cell(const Items &I, long long l)
{
    const long long c0 = l % I.ld0, q = l / I.ld0, c1 = q % I.ld1, c2 = q / I.ld1;
    return (c0 / I.T0) + (long long)I.nt0 * ((c1 / I.T1) + (long long)I.nt1 * (c2 / I.T2));
}
// record of every listed item (one thread per item, build time)
__global__ void kf_tile_records(Items I, TileRec *rec)
{
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < I.n; it += gridDim.x * blockDim.x) {
        const unsigned v = (unsigned)I.it[it];
        const int f = (int)(v >> 30);
        const long long ord = (long long)(v & 0x3fffffffu);
        TileRec R;
        R.f = (signed char)f; R.full = 0; R.ghost = 0; R.pad_ = 0; R.ox = R.oy = R.oz = 0;
        if (f >= 2) {
            R.base = (long long)I.wlo + ord * FTILE;
            R.baseq = R.base;
            const long long left = (long long)I.whi - R.base;
            R.nx = (short)(left < FTILE ? left : FTILE);
            R.ylo = 0; R.yhi = 1; R.zlo = 0; R.zhi = 1;
        } else {
            const long long t0 = ord % I.nt0, q = ord / I.nt0, t1 = q % I.nt1, t2 = q / I.nt1;
            const long long o[3] = {t0 * I.T0, t1 * I.T1, t2 * I.T2};
            const long long ld[3] = {I.ld0, I.ld1, I.ld2};
            const int T[3] = {I.T0, I.T1, I.T2};
            int lo[3], hi[3];
            for (int d = 0; d < 3; ++d) {
                long long a = 0, b = ld[d] - o[d];
                if (d == I.sd) { a = 1 - o[d]; b = (long long)I.lz - 1 - o[d]; }   // owned planes 1 .. lz-2 of the slab dimension
                if (a < 0) a = 0;
                if (b > T[d]) b = T[d];
                if (b < a) b = a;
                lo[d] = (int)a; hi[d] = (int)b;
            }
            R.base = o[0] + I.ld0 * (o[1] + I.ld1 * o[2]);
            R.baseq = o[0] + I.P0 * (o[1] + I.ld1 * o[2]);
            R.ox = (int)o[0]; R.oy = (int)o[1]; R.oz = (int)o[2];
            // x validity is [0, nx): the slab dimension is x only for 1-D grids, where the lower ghost is excluded by shifting the base
            if (I.sd == 0) { R.base += lo[0]; R.baseq += lo[0]; R.nx = (short)(hi[0] - lo[0]); }
            else R.nx = (short)hi[0];
            // does the box (tile + one halo cell) reach a ghost plane of the slab dimension that a neighbour rank fills?
            if (I.sd > 0) {
                const long long osd = o[I.sd];
                const int Tsd = T[I.sd];
                if ((I.glo && osd - 1 <= 0) || (I.ghi && osd + Tsd >= I.lz - 1)) R.ghost = 1;
            }
            R.ylo = (signed char)lo[1]; R.yhi = (signed char)hi[1]; R.zlo = (signed char)lo[2]; R.zhi = (signed char)hi[2];
            R.full = (lo[0] == 0 && hi[0] == T[0] && lo[1] == 0 && hi[1] == T[1] && lo[2] == 0 && hi[2] == T[2]) ? 1 : 0;
        }
        rec[it] = R;
    }
}

// result slots of the folded Krylov loops (dense-part and band-part partial sums are adjacent: one allreduce covers both)
// (rho, rr, rho_band, rho_poly_band) groups live in two ping-pong sets {0..3} and {4..7}: one allreduce of 4 doubles publishes a new group
// FS_SIG_G: (p, v) partial of the ghost-class tiles (fold2.cuh); FS_ALPHA / FS_XPEND: alpha_k and "x += alpha_k p_k still pending" of the fused iteration
enum { FS_PAIR0 = 0, FS_PAIR1 = 4, FS_SIG_D = 8, FS_SIG_B = 9, FS_SIG_G = 10, FS_TS_D = 11, FS_TT_D = 12, FS_TS_B = 13, FS_TT_B = 14, FS_BB = 15, FS_RR0 = 16, FS_ITERS = 17,
       FS_TMP = 18 /* and 19 */, FS_ALPHA = 20, FS_XPEND = 21 };
static_assert(FS_XPEND < RED_SLOTS, "result slots");
#define FS_TRIPLE(p) (4 * (p))
#define FS_NGROUP 4
// rho of the group at slot sl = (r, z) with z = q(M^) r + (q_B(M^_BB) - 1) r_B:
//   [sl]     (r, q(M^) r), dense part of the last polynomial step -- (r, r) without the polynomial preconditioner (q = 1),
//   [sl + 2] (r_B, z_B - r_B), the band preconditioner's part (kf_band_poly),
//   [sl + 3] the band kernel's share of (r, q(M^) r) (0 without the polynomial);           [sl + 1] is ||r||^2 (stopping test)
__device__ __forceinline__ double rho_at(const double *res, int sl) { return res[sl] + res[sl + 2] + res[sl + 3]; }

// device-side stopping test: the Krylov kernels of an iteration turn into no-ops once ||r||^2 (slot `sl_rr`) is below the
// tolerance, so the host may queue several iterations between two looks at the residual without doing extra work
struct StopCrit { double rtol2, atol2; int sl_rr; };
__device__ __forceinline__ bool fold_done(const double *res, const StopCrit &sc)
{
    const double rr = res[sc.sl_rr];
    return rr <= fmax(sc.rtol2 * res[FS_BB], sc.atol2) || !(rr == rr);
}

// ---- face coefficients ---------------------------------------------------------------------------------------------------
// lower face of cell l in direction d (id = coordinate of l along d): W!, (a, cw) on (u_l, w_l), (b, dw) on (u_{l-s}, w_{l-s})
__device__ __forceinline__ void face_coef(const PhaseDev &ph, const Grid &g, long long l, int d, int id, double kap, double &W, double &a, double &cw,
                                          double &b, double &dw)
{
    const int last = g.pd[d] - 1;
    const double ei = id < last ? 1.0 : 0.0;
    const double Al = ph.A[d][l], Bl = ph.B[d][l];
    W = ph.Wd[d][l];
    a = ei * Bl;
    cw = kap * ei * (Al - Bl);
    if (id > 0) { const double Bm = ph.B[d][l - g.stride[d]]; b = -Bm; dw = -kap * (Al - Bm); }
    else { b = 0.0; dw = 0.0; }
}

// per-cell diagonal block of M restricted to the active unknowns: D[0]=D00 D[1]=D11 D[2]=D02 D[3]=D12 D[4]=D22
template <int N>
__device__ __forceinline__ void cell_block(const FoldDev &fd, const Grid &g, long long l, const int c[PB_MAXD], bool act0, bool act1, bool actw, double D[5])
{
    D[0] = D[1] = D[2] = D[3] = D[4] = 0.0;
    for (int p = 0; p < fd.nbulk; ++p) {
        const PhaseDev &ph = fd.ph[p];
        const double sp = fd.s[p];
        double Dpp = sp * fd.cVc * ph.V[l] / D_at(ph, l), Dpw = 0.0, Dww = 0.0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            double W, a, cw, b, dw;
            face_coef(ph, g, l, d, c[d], fd.kap[p], W, a, cw, b, dw);
            Dpp += sp * W * a * a; Dpw += sp * W * a * cw; Dww += sp * W * cw * cw;
            if (c[d] < g.pd[d] - 1) {
                face_coef(ph, g, l + g.stride[d], d, c[d] + 1, fd.kap[p], W, a, cw, b, dw);
                Dpp += sp * W * b * b; Dpw += sp * W * b * dw; Dww += sp * W * dw * dw;
            }
        }
        D[p] = Dpp; D[2 + p] = Dpw; D[4] += Dww;
    }
    D[4] += fd.mwc * fd.ph[0].Gam[l];
    if (!act0) { D[0] = 1.0; D[2] = 0.0; }
    if (!act1) { D[1] = 1.0; D[3] = 0.0; }
    if (!actw) { D[4] = 1.0; D[2] = 0.0; D[3] = 0.0; }
}

// L^-1 of a cell: (i00, i11, i20, i21, i22); diagonal (sc0, sc1, 0) outside the band
__device__ __forceinline__ void cell_linv(const FoldDev &fd, long long l, double I[5])
{
    const int bo = fd.bord ? fd.bord[l] : -1;
    if (bo >= 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) I[k] = fd.Linv[(size_t)k * fd.nB + bo];
    } else {
        I[0] = fd.sc[0][l]; I[1] = fd.nbulk > 1 ? fd.sc[1][l] : 0.0; I[2] = I[3] = I[4] = 0.0;
    }
}
// R = Li * O * Lj^T for lower-triangular Li, Lj given as (i00,i11,i20,i21,i22)
__device__ __forceinline__ void tri_sandwich(const double Li[5], const double O[9], const double Lj[5], double R[9])
{
    double T[9];   // T = Li * O
#pragma unroll
    for (int cidx = 0; cidx < 3; ++cidx) {
        T[0 * 3 + cidx] = Li[0] * O[0 * 3 + cidx];
        T[1 * 3 + cidx] = Li[1] * O[1 * 3 + cidx];
        T[2 * 3 + cidx] = Li[2] * O[0 * 3 + cidx] + Li[3] * O[1 * 3 + cidx] + Li[4] * O[2 * 3 + cidx];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {   // R = T * Lj^T : R[r][0] = T[r][0] Lj00 ; R[r][1] = T[r][1] Lj11 ; R[r][2] = T[r][0] Lj20 + T[r][1] Lj21 + T[r][2] Lj22
        R[r * 3 + 0] = T[r * 3 + 0] * Lj[0];
        R[r * 3 + 1] = T[r * 3 + 1] * Lj[1];
        R[r * 3 + 2] = T[r * 3 + 0] * Lj[2] + T[r * 3 + 1] * Lj[3] + T[r * 3 + 2] * Lj[4];
    }
}

// ---- set-up kernels -------------------------------------------------------------------------------------------------------
__global__ void kf_mark_band(long long nloc, const unsigned char *__restrict__ mw, long long *list, int *count, int cap)
{
    for (long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x; l < nloc; l += (long long)gridDim.x * blockDim.x)
        if (mw[l] & MB_IFREE) { const int k = atomicAdd(count, 1); if (k < cap) list[k] = l; }
}
__global__ void kf_bord_fill(int nB, const long long *__restrict__ Bcell, int *__restrict__ bord)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nB; k += gridDim.x * blockDim.x) bord[Bcell[k]] = k;
}
__global__ void kf_eord_fill(int nE, const long long *__restrict__ Ecell, const int *__restrict__ bord, int *__restrict__ eord, int *__restrict__ EofB)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
        const long long l = Ecell[e];
        eord[l] = e;
        const int bo = bord[l];
        if (bo >= 0) EofB[bo] = e;
    }
}
// (after kf_blocks: EnbrB is known) E index of the band neighbours
__global__ void kf_enbre_fill(Grid g, int nE, int nEp, const long long *__restrict__ Ecell, const int *__restrict__ EnbrB, const int *__restrict__ eord, int *__restrict__ EnbrE)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nE; e += gridDim.x * blockDim.x) {
        const long long l = Ecell[e];
        for (int kk = 0; kk < 2 * g.N; ++kk) {
            const int nb = EnbrB[(size_t)kk * nEp + e];
            const long long ln = (kk & 1) ? l + g.stride[kk >> 1] : l - g.stride[kk >> 1];
            EnbrE[(size_t)kk * nEp + e] = nb >= 0 ? eord[ln] : -1;
        }
    }
}

template <int N>
__global__ void kf_diag(Grid g, FoldDev fd)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nown; t += (long long)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const long long l = t + g.plane;
        const bool act0 = fd.m[0][l] & MB_FREE, act1 = fd.nbulk > 1 && (fd.m[1][l] & MB_FREE), actw = fd.has_w && (fd.mw[l] & MB_IFREE);
        if (!act0 && !act1 && !actw) { fd.sc[0][l] = 0.0; if (fd.nbulk > 1) fd.sc[1][l] = 0.0; continue; }
        double D[5];
        cell_block<N>(fd, g, l, c, act0, act1, actw, D);
        if (!actw) {
            fd.sc[0][l] = (act0 && D[0] > 0.0) ? 1.0 / sqrt(D[0]) : 0.0;
            if (fd.nbulk > 1) fd.sc[1][l] = (act1 && D[1] > 0.0) ? 1.0 / sqrt(D[1]) : 0.0;
        } else {
            fd.sc[0][l] = 0.0;
            if (fd.nbulk > 1) fd.sc[1][l] = 0.0;
            const int bo = fd.bord[l];
            const double l00 = sqrt(D[0]), l11 = sqrt(D[1]);
            const double l20 = D[2] / l00, l21 = D[3] / l11;
            double S = D[4] - l20 * l20 - l21 * l21;
            if (!(S > 1e-14 * D[4])) S = 1e-14 * D[4];
            const double l22 = sqrt(S);
            const double i00 = act0 ? 1.0 / l00 : 0.0, i11 = act1 ? 1.0 / l11 : 0.0, i22 = 1.0 / l22;
            fd.Linv[(size_t)0 * fd.nB + bo] = i00;
            fd.Linv[(size_t)1 * fd.nB + bo] = i11;
            fd.Linv[(size_t)2 * fd.nB + bo] = -l20 * i00 * i22;
            fd.Linv[(size_t)3 * fd.nB + bo] = -l21 * i11 * i22;
            fd.Linv[(size_t)4 * fd.nB + bo] = i22;
        }
    }
}

template <int N>
__global__ void kf_off(Grid g, FoldDev fd)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nown; t += (long long)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const long long l = t + g.plane;
        const bool band = fd.bord && fd.bord[l] >= 0;
        for (int p = 0; p < fd.nbulk; ++p) {
            const double sl = fd.sc[p][l];
#pragma unroll
            for (int d = 0; d < N; ++d) {
                double v = 0.0;
                if (sl != 0.0 && c[d] > 0 && !band) {
                    const long long ln = l - g.stride[d];
                    const double sn = fd.sc[p][ln];
                    if (sn != 0.0 && !(fd.bord && fd.bord[ln] >= 0)) {
                        double W, a, cw, b, dw;
                        face_coef(fd.ph[p], g, l, d, c[d], fd.kap[p], W, a, cw, b, dw);
                        v = fd.s[p] * W * a * b * sl * sn;
                    }
                }
                fd.off[p][d][l] = v;
            }
        }
    }
}

template <int N>
__global__ void kf_mark_E(Grid g, const int *__restrict__ bord, long long *list, int *count, int cap)
{
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < g.nown; t += (long long)gridDim.x * blockDim.x) {
        int c[PB_MAXD];
        cell_coords(g, t, c);
        const long long l = t + g.plane;
        bool in = bord[l] >= 0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            if (c[d] > 0) in = in || bord[l - g.stride[d]] >= 0;
            if (c[d] < g.pd[d] - 1) in = in || bord[l + g.stride[d]] >= 0;
        }
        if (in) { const int k = atomicAdd(count, 1); if (k < cap) list[k] = l; }
    }
}

// coupling block between cell l (rows) and its neighbour across the lower face of `lf` (lf = l: neighbour l-s; lf = l+s: neighbour l+s)
template <int N>
__device__ __forceinline__ void face_block(const FoldDev &fd, const Grid &g, long long lf, int d, int idf, bool rows_are_upper, double O[9])
{
#pragma unroll
    for (int k = 0; k < 9; ++k) O[k] = 0.0;
    for (int p = 0; p < fd.nbulk; ++p) {
        double W, a, cw, b, dw;
        face_coef(fd.ph[p], g, lf, d, idf, fd.kap[p], W, a, cw, b, dw);
        const double sw = fd.s[p] * W;
        // upper cell (lf) vector: (a on comp p, cw on comp 2); lower cell (lf - s) vector: (b on comp p, dw on comp 2)
        const double ru[2] = {a, cw}, rl[2] = {b, dw};
        const double *rr = rows_are_upper ? ru : rl, *cc = rows_are_upper ? rl : ru;
        O[p * 3 + p] += sw * rr[0] * cc[0];
        O[p * 3 + 2] += sw * rr[0] * cc[1];
        O[2 * 3 + p] += sw * rr[1] * cc[0];
        O[2 * 3 + 2] += sw * rr[1] * cc[1];
    }
}

template <int N>
__global__ void kf_blocks(Grid g, FoldDev fd)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < fd.nE; e += gridDim.x * blockDim.x) {
        const long long l = fd.Ecell[e];
        int c[PB_MAXD];
        cell_coords(g, l - g.plane, c);
        const int bo = fd.bord[l];
        fd.EB[e] = bo;
        double Li[5], O[9], R[9];
        cell_linv(fd, l, Li);
        // self block: L^-1 D L^-T - I (exactly what the scaled operator carries on the diagonal block, no cancellation assumed)
        if (bo >= 0) {
            const bool act0 = fd.m[0][l] & MB_FREE, act1 = fd.nbulk > 1 && (fd.m[1][l] & MB_FREE);
            double D[5];
            cell_block<N>(fd, g, l, c, act0, act1, true, D);
            O[0] = act0 ? D[0] : 0.0; O[1] = 0.0; O[2] = D[2];
            O[3] = 0.0; O[4] = act1 ? D[1] : 0.0; O[5] = D[3];
            O[6] = D[2]; O[7] = D[3]; O[8] = D[4];
            tri_sandwich(Li, O, Li, R);
            if (act0) R[0] -= 1.0;
            if (act1) R[4] -= 1.0;
            R[8] -= 1.0;
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = 0.0;
        }
        double *__restrict__ eb = fd.Eblk + e;   // SoA: entry j of cell e at eb[j * nEp]
        const size_t ES = (size_t)fd.nEp;
#pragma unroll
        for (int k = 0; k < 9; ++k) eb[k * ES] = R[k];
        // cells whose tile the fused kernel processes (interior class): local plane of the slab dimension in [fix_lo, fix_hi)
        unsigned fixm = 0;
        {
            const int pl = (int)(l / g.plane);
            if (pl >= fd.fix_lo && pl < fd.fix_hi) fixm |= 1u;
        }
#pragma unroll
        for (int d = 0; d < N; ++d) {
            const long long s = g.stride[d];
            int nbL = -1, nbU = -1;
            double RL[9], RU[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { RL[k] = 0.0; RU[k] = 0.0; }
            if (c[d] > 0) {
                nbL = fd.bord[l - s];
                if (bo >= 0 || nbL >= 0) {
                    double Lj[5];
                    cell_linv(fd, l - s, Lj);
                    face_block<N>(fd, g, l, d, c[d], true, O);
                    tri_sandwich(Li, O, Lj, RL);
                }
            }
            if (c[d] < g.pd[d] - 1) {
                nbU = fd.bord[l + s];
                if (bo >= 0 || nbU >= 0) {
                    double Lj[5];
                    cell_linv(fd, l + s, Lj);
                    face_block<N>(fd, g, l + s, d, c[d] + 1, false, O);
                    tri_sandwich(Li, O, Lj, RU);
                }
            }
            fd.EnbrB[(size_t)(2 * d) * ES + e] = nbL;
            fd.EnbrB[(size_t)(2 * d + 1) * ES + e] = nbU;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                eb[((1 + 2 * d) * 9 + k) * ES] = RL[k];
                eb[((2 + 2 * d) * 9 + k) * ES] = RU[k];
            }
            {
                const int pl = (int)(l / g.plane);
                const int plL = d == g.sd ? pl - 1 : pl, plU = d == g.sd ? pl + 1 : pl;
                if (plL >= fd.fix_lo && plL < fd.fix_hi) fixm |= 1u << (1 + 2 * d);
                if (plU >= fd.fix_lo && plU < fd.fix_hi) fixm |= 1u << (2 + 2 * d);
            }
        }
        fd.Efix[e] = (unsigned char)fixm;
    }
}

// active tile census: flags[f * ntile + tile] = 1 if any cell of the tile carries a free unknown of bulk field f
__global__ void kf_tile_flags(Grid g, Items I, int nbulk, const unsigned char *__restrict__ m0, const unsigned char *__restrict__ m1, long long ntile, int *flags)
{
    for (long long l = g.plane + blockIdx.x * (long long)blockDim.x + threadIdx.x; l < g.plane + g.nown; l += (long long)gridDim.x * blockDim.x) {
        const bool a0 = m0[l] & MB_FREE, a1 = nbulk > 1 && (m1[l] & MB_FREE);
        if (a0 || a1) {
            const long long t = tile_of_cell(I, l);
            if (a0) flags[t] = 1;
            if (a1) flags[ntile + t] = 1;
        }
    }
}
__global__ void kf_tile_list(int nbulk, long long ntile, const int *__restrict__ flags, int wchunks, int *items, int *count)
{
    const long long tot = (long long)nbulk * ntile + wchunks;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        int v = -1;
        if (i < (long long)nbulk * ntile) { if (flags[i]) v = (int)(((unsigned)(i / ntile) << 30) | (unsigned)(i % ntile)); }
        else v = (int)((2u << 30) | (unsigned)(i - (long long)nbulk * ntile));
        if (v != -1) items[atomicAdd(count, 1)] = v;
    }
}
// per-item coefficient census: a tile whose coefficients off_d[l] and off_d[l + s_d] are one constant per direction over all its
// cells (every tile of full cells away from the interface and the border) is applied without reading the coefficient arrays.
// Also counts the cells of uniform / general tiles (results[0], results[1]) for the roofline accounting.
template <int N>
__global__ void __launch_bounds__(FCH) kf_tile_meta(Grid g, FoldDev fd, Items I, unsigned char *uni, double *ucoef, double utol, double *partials, double *results,
                                                    unsigned *counter)
{
    __shared__ double smn[PB_MAXD][FCH / 32], smx[PB_MAXD][FCH / 32];
    __shared__ int s_uni;
    double cnt[3] = {0.0, 0.0, 0.0};   // cells of constant-coefficient tiles, of streamed-coefficient tiles, of the staged interior branch
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        if (R.f >= 2) { if (threadIdx.x == 0) uni[it] = 0; continue; }   // uniform over the block
        double mn[PB_MAXD], mx[PB_MAXD];
        int nok = 0, hasb = 0, hase = 0;
#pragma unroll
        for (int d = 0; d < N; ++d) { mn[d] = 1e300; mx[d] = -1e300; }
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            long long l;
            if (!tile_cell(I, R, k, l)) continue;
            ++nok;
            if (fd.bord && fd.bord[l] >= 0) hasb = 1;
            if (fd.eord && fd.eord[l] >= 0) hase = 1;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                const double *__restrict__ of = R.f == 0 ? fd.off[0][d] : fd.off[1][d];
                const double a = of[l], b = of[l + g.stride[d]];
                mn[d] = fmin(mn[d], fmin(a, b)); mx[d] = fmax(mx[d], fmax(a, b));
            }
        }
#pragma unroll
        for (int d = 0; d < N; ++d) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
            if ((threadIdx.x & 31) == 0) { smn[d][threadIdx.x >> 5] = mn[d]; smx[d][threadIdx.x >> 5] = mx[d]; }
        }
        const int anyb = __syncthreads_or(hasb);
        const int anye = __syncthreads_or(hase);
        if (threadIdx.x == 0) {
            bool u = true;
            for (int d = 0; d < N; ++d) {
                double a = 1e300, b = -1e300;
                for (int w = 0; w < FCH / 32; ++w) { a = fmin(a, smn[d][w]); b = fmax(b, smx[d][w]); }
                // equal up to round-off of the geometry: on grids whose spacing is not a power of two the capacities of full cells carry the
                // cancellation error of h_j = x_{j+1} - x_j (~1e-14 relative), and so do the folded coefficients.  Tiles whose coefficients
                // agree to 1e-12 relative use the mid value -- three orders below the 1e-9 parity bar; utol = 0 demands bitwise equality.
                u = u && (b - a <= utol * fmax(fabs(a), fabs(b)));
                ucoef[(size_t)it * PB_MAXD + d] = 0.5 * (a + b);
            }
            uni[it] = (unsigned char)((u ? 1 : 0) | (anyb ? 2 : 0) | (anye ? 8 : 0));   // bit 0: constant coefficients, bit 1: holds band cells, bit 3: holds band or fringe cells (E list)
            s_uni = u ? 1 : 0;
        }
        __syncthreads();
        cnt[s_uni ? 0 : 1] += (double)nok;
        if (s_uni && R.full && !anyb) cnt[2] += (double)nok;
        __syncthreads();
    }
    block_reduce_publish<3>(cnt, partials, results, counter);
}

// ---- operator ---------------------------------------------------------------------------------------------------------------
// dense part: y = x + sum_d off_d[l] x[l-s] + off_d[l+s] x[l+s] on the active chunks of each bulk field.
// MODE 0: no dot; 1: publish (x, y); 2: publish (aux, y); 3: publish (y, x), (y, y);
// MODE 4: one step of the polynomial preconditioner, y = pc.r aux + pc.z x + pc.A (M^ x), publish (aux, y)   (aux = the residual)
struct PolyCoef { double r, z, A; };
template <int N, int MODE>
__global__ void __launch_bounds__(FCH, 4) kf_apply_dense(Grid g, FoldDev fd, Items I, FVec x, FVec y, FVec aux, double *partials, double *results, unsigned *counter,
                                                      const double *res, StopCrit stop, PolyCoef pc)
{
    if (stop.sl_rr >= 0 && fold_done(res, stop)) return;
    double v[2] = {0.0, 0.0};
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        // the record, flags and constants of the block's NEXT item are pulled into L1 while this one streams: the dependent chain
        // record -> addresses -> data then costs an L1 hit per tile instead of an L2 / HBM round trip
        if (threadIdx.x == 0 && it + (int)gridDim.x < I.n) {
            const int nx = it + gridDim.x;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(I.rec + nx));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(I.uni + nx));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(I.ucoef + (size_t)nx * PB_MAXD));
        }
        const TileRec R = I.rec[it];
        const bool uni = (I.uni[it] & 1) != 0;
        if (R.f >= 2) continue;
        const int f = R.f;
        const double *__restrict__ xf = f == 0 ? x.f[0] : x.f[1];
        double *__restrict__ yf = f == 0 ? y.f[0] : y.f[1];
        const double *__restrict__ af = f == 0 ? aux.f[0] : aux.f[1];
        if (N >= 2 && uni && R.full) {
            // interior tile of full cells: constant coefficients, every cell valid.  The thread owns FU cells that are consecutive along
            // the k direction (y in 2-D, z in 3-D): their k-neighbours are shared registers (FU + 2 loads for the column), the x-neighbours
            // come from the adjacent lanes (a warp is one 32-cell row; only lanes 0 and 31 load the halo), y-neighbours are loaded in 3-D.
            const double *__restrict__ uc = I.ucoef + (size_t)it * PB_MAXD;
            const int lane = (int)threadIdx.x & 31, ty = (int)threadIdx.x >> 5;
            const long long ks = I.ustrideq;                       // (Krylov vectors: re-pitched layout)
            const long long l0 = R.baseq + lane + (long long)(ty * I.tym) * I.P0;
            const double *__restrict__ p0 = xf + l0;
            double col[FU + 2], yn[FU][2], avv[FU];
#pragma unroll
            for (int k = -1; k <= FU; ++k) col[k + 1] = p0[k * ks];
            if (N == 3) {
#pragma unroll
                for (int k = 0; k < FU; ++k) { yn[k][0] = p0[k * ks - I.P0]; yn[k][1] = p0[k * ks + I.P0]; }
            }
            if (MODE == 2 || MODE == 4) {
#pragma unroll
                for (int k = 0; k < FU; ++k) avv[k] = af[l0 + k * ks];
            }
            double lf[FU], rt[FU];
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                lf[k] = __shfl_up_sync(0xffffffffu, col[k + 1], 1);
                rt[k] = __shfl_down_sync(0xffffffffu, col[k + 1], 1);
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < FU; ++k) lf[k] = p0[k * ks - 1];
            }
            if (lane == 31) {
#pragma unroll
                for (int k = 0; k < FU; ++k) rt[k] = p0[k * ks + 1];
            }
            const double cx = uc[0], ck = uc[N - 1], cy = uc[1];
#pragma unroll
            for (int k = 0; k < FU; ++k) {
                double acc = col[k + 1] + cx * (lf[k] + rt[k]) + ck * (col[k] + col[k + 2]);
                if (N == 3) acc += cy * (yn[k][0] + yn[k][1]);
                if (MODE == 4) acc = pc.r * avv[k] + pc.z * col[k + 1] + pc.A * acc;
                yf[l0 + k * ks] = acc;
                if (MODE == 1) v[0] += col[k + 1] * acc;
                if (MODE == 2 || MODE == 4) v[0] += avv[k] * acc;
                if (MODE == 3) { v[0] += acc * col[k + 1]; v[1] += acc * acc; }
            }
            continue;
        }
        // general tiles (interface band, domain border ring, partially owned tiles): coefficients streamed (or the tile constants when
        // only the validity is partial).  GP cells per thread and round, every load of the round unconditional so that all of them are in
        // flight together: an invalid cell loads from a cell of the tile that is certainly valid (tile-relative (0, ylo, zlo)) and only
        // its store and its share of the dot product are predicated.  (One cell at a time -- four dependent round trips per tile -- made
        // these tiles, ~12 % of the cells at 2048^2 and 28 % at 384^3, the larger part of the kernel's time.)
        const double *__restrict__ uc = I.ucoef + (size_t)it * PB_MAXD;
        constexpr int GP = PB_APPLY_GP(N);
        const long long lsafe = R.base + (long long)R.ylo * I.ld0 + (long long)R.zlo * I.ld0 * I.ld1;
        const long long qsafe = R.baseq + (long long)R.ylo * I.P0 + (long long)R.zlo * I.P0 * I.ld1;
        const double *__restrict__ of0 = f == 0 ? fd.off[0][0] : fd.off[1][0];
        const double *__restrict__ of1 = f == 0 ? fd.off[0][N > 1 ? 1 : 0] : fd.off[1][N > 1 ? 1 : 0];
        const double *__restrict__ of2 = f == 0 ? fd.off[0][N > 2 ? 2 : 0] : fd.off[1][N > 2 ? 2 : 0];
#pragma unroll 1
        for (int k0 = 0; k0 < FU; k0 += GP) {
            long long l[GP], lq[GP];
            bool ok[GP];
            double xl[GP], av[GP], cm[GP][N], cp[GP][N], xm[GP][N], xp[GP][N];
#pragma unroll
            for (int j = 0; j < GP; ++j) {
                long long t, tq;
                ok[j] = tile_cell(I, R, k0 + j, t, tq);
                l[j] = ok[j] ? t : lsafe;
                lq[j] = ok[j] ? tq : qsafe;
            }
#pragma unroll
            for (int j = 0; j < GP; ++j) {
                xl[j] = xf[lq[j]];
                av[j] = (MODE == 2 || MODE == 4) ? af[lq[j]] : 0.0;
#pragma unroll
                for (int d = 0; d < N; ++d) {
                    const long long s = g.stride[d], sq = fd.sq[d];
                    const double *__restrict__ of = d == 0 ? of0 : (d == 1 ? of1 : of2);
                    if (uni) { cm[j][d] = uc[d]; cp[j][d] = uc[d]; }
                    else { cm[j][d] = of[l[j]]; cp[j][d] = of[l[j] + s]; }
                    xm[j][d] = xf[lq[j] - sq];
                    xp[j][d] = xf[lq[j] + sq];
                }
            }
#pragma unroll
            for (int j = 0; j < GP; ++j) {
                double acc = xl[j];
#pragma unroll
                for (int d = 0; d < N; ++d) acc += cm[j][d] * xm[j][d] + cp[j][d] * xp[j][d];
                if (MODE == 4) acc = pc.r * av[j] + pc.z * xl[j] + pc.A * acc;
                if (ok[j]) {
                    yf[lq[j]] = acc;
                    if (MODE == 1) v[0] += xl[j] * acc;
                    if (MODE == 2 || MODE == 4) v[0] += av[j] * acc;
                    if (MODE == 3) { v[0] += acc * xl[j]; v[1] += acc * acc; }
                }
            }
        }
    }
    if (MODE == 1 || MODE == 2 || MODE == 4) { double w[1] = {v[0]}; block_reduce_publish<1>(w, partials, results, counter); }
    if (MODE == 3) block_reduce_publish<2>(v, partials, results, counter);
}

// One THREAD per band / fringe cell e (consecutive threads on consecutive cells of the sorted list): the (1 + 2N) 3 x 3 coefficient blocks are
// stored SoA ([entry][cell]: coalesced), the 3 unknowns of the cell and of its 2N neighbours are gathered by the thread itself -- up to
// 9 (1 + 2N) + 3 (1 + 2N) independent loads in flight per thread instead of 2-4 per lane of the former warp-per-cell kernel (measured:
// the band kernels were 30 % of a 3-D diphasic iteration, 4-6 x off their bandwidth).  Blocks between two non-band cells are zero by
// construction (the dense stencil carries those couplings) and are skipped.
// BAND_ONLY restricts the columns to band cells (the principal submatrix M^_BB) and adds the identity diagonal.
// dzadd (fused iteration): the bulk components of band cells are gathered as x + dz where Efix says so (see kf_apply_band).
// LPC lanes share one cell (lane `sub` takes the blocks k = sub, sub + LPC, ...): 1 for large bands (3-D: every SM is busy with one thread per cell,
// and a thread keeps up to 84 independent loads in flight), 8 for small ones (2-D: a few ten thousand cells -- shorter chains, 8 x the threads).
template <int N, bool BAND_ONLY, int LPC>
__device__ __forceinline__ void band_rows(const Grid &g, const FoldDev &fd, int e, const FVec &x, double &a0, double &a1, double &a2, double &x0,
                                          double &x1, double &xw, long long &l, int &bo, int sub, const double *__restrict__ dzadd = nullptr)
{
    (void)g;
    const size_t ES = (size_t)fd.nEp;
    l = fd.Ecell[e];
    bo = fd.EB[e];
    const long long lq = fd_q(fd, l);
    const bool two = fd.nbulk > 1;
    const unsigned fixm = dzadd != nullptr ? fd.Efix[e] : 0u;
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    x0 = x1 = xw = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 1 + 2 * N; k0 += LPC) {
        const int k = k0 + sub;
        if (LPC > 1 && k >= 1 + 2 * N) break;
        long long ln = lq;
        int nb = bo;
        if (k > 0) {
            const int kk = k - 1, d = kk >> 1;
            ln = (kk & 1) ? lq + fd.sq[d] : lq - fd.sq[d];
            nb = fd.EnbrB[(size_t)kk * ES + e];
        }
        // BAND_ONLY: the preconditioner block is the band block of THIS rank (neighbours in the ghost planes are left out), so that
        // applying it needs no halo exchange; the operator itself (BAND_ONLY == false) couples across ranks as usual
        const bool use = !(BAND_ONLY && (nb < 0 || (k > 0 && (nb < fd.nBlo || nb >= fd.nBlo + fd.nBown))));
        const bool nonzero = k == 0 ? bo >= 0 : (bo >= 0 || nb >= 0);     // (kf_blocks: other blocks are exactly zero)
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (use && (nonzero || k == 0)) {
            v0 = x.f[0][ln];
            if (two) v1 = x.f[1][ln];
            if (nb >= 0) v2 = x.f[2][nb];
            if (nb >= 0 && ((fixm >> k) & 1u)) { v0 += dzadd[nb]; if (two) v1 += dzadd[(size_t)fd.nB + nb]; }
        }
        if (k == 0) { x0 = v0; x1 = v1; xw = v2; }
        if (use && nonzero) {
            const double *__restrict__ c = fd.Eblk + (size_t)(k * 9) * ES + e;
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
        x0 = __shfl_sync(0xffffffffu, x0, 0, LPC); x1 = __shfl_sync(0xffffffffu, x1, 0, LPC); xw = __shfl_sync(0xffffffffu, xw, 0, LPC);
    }
    a0 = r0; a1 = r1; a2 = r2;
    if (BAND_ONLY) { a0 += x0; a1 += x1; a2 += xw; }
}
// cell loop shared by the band kernels: groups of LPC lanes walk the cell list; groups past the end redo the last cell (the shuffles are warp-wide)
#define BAND_LOOP(LPC_)                                                                              \
    const int sub = (int)threadIdx.x % (LPC_), gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) / (LPC_)); \
    const int ngroups = (int)(gridDim.x * blockDim.x / (LPC_));                                      \
    const int rounds = fd.nE > 0 ? (fd.nE + ngroups - 1) / ngroups : 0;                              \
    for (int rd = 0; rd < rounds; ++rd)                                                              \
        if (const int e_raw = rd * ngroups + gid; true)                                              \
            if (const bool live = e_raw < fd.nE; true)                                               \
                if (const int e = live ? e_raw : fd.nE - 1; true)

// band part (after the dense kernel): adds every coupling that involves a band cell; w rows are written here.
// dzfix (MODE 1, fused iteration): x holds p' = r + beta p_old on the bulk unknowns of the band cells of the fused kernel's tiles; the correction dz
// is added HERE -- to the gathered values and to y (the dense kernel wrote y = p' on band cells: their dense couplings are all zero); x itself is
// corrected by kf_band_poly (other threads still gather p' here) -- so that the tile kernels need no band look-ups at all.
template <int N, int MODE, int LPC>
__global__ void __launch_bounds__(256) kf_apply_band(Grid g, FoldDev fd, FVec x, FVec y, FVec aux, double *partials, double *results, unsigned *counter,
                                                     const double *res, StopCrit stop, PolyCoef pc, const double *__restrict__ dzfix = nullptr)
{
    if (stop.sl_rr >= 0 && fold_done(res, stop)) return;
    double v[2] = {0.0, 0.0};
    const bool two = fd.nbulk > 1;
    BAND_LOOP(LPC) {
        double a0, a1, a2, x0, x1, xw;
        long long l; int bo;
        band_rows<N, false, LPC>(g, fd, e, x, a0, a1, a2, x0, x1, xw, l, bo, sub, dzfix);
        if (sub != 0 || !live) continue;
        const long long lq = fd_q(fd, l);
        const double y0p = y.f[0][lq], y1p = two ? y.f[1][lq] : 0.0;
        if (MODE == 4) { a0 *= pc.A; a1 *= pc.A; }   // the dense kernel has written pc.r aux + pc.z x + pc.A (dense part of M^ x)
        double d0 = 0.0, d1 = 0.0;
        if (MODE == 1 && dzfix != nullptr && bo >= fd.nBlo && bo < fd.nBlo + fd.nBown && (fd.Efix[e] & 1u)) {   // band cell of a fused-kernel tile: y = p' -> p_k (+ couplings below)
            d0 = dzfix[bo]; d1 = two ? dzfix[(size_t)fd.nB + bo] : 0.0;
            v[0] += d0 * y0p + d1 * y1p + x0 * d0 + x1 * d1;
        }
        y.f[0][lq] = y0p + d0 + a0;
        if (two) y.f[1][lq] = y1p + d1 + a1;
        double yw = 0.0;
        if (bo >= 0) {
            yw = xw + a2;
            if (MODE == 4) yw = pc.r * aux.f[2][bo] + pc.z * xw + pc.A * yw;
            y.f[2][bo] = yw;
        }
        if (MODE == 1) v[0] += x0 * a0 + x1 * a1 + xw * yw;
        if (MODE == 2 || MODE == 4) v[0] += aux.f[0][lq] * a0 + (two ? aux.f[1][lq] * a1 : 0.0) + (bo >= 0 ? aux.f[2][bo] * yw : 0.0);
        if (MODE == 3) {
            v[0] += x0 * a0 + x1 * a1 + xw * yw;
            v[1] += (2.0 * y0p + a0) * a0 + (2.0 * y1p + a1) * a1 + yw * yw;
        }
    }
    if (MODE == 1 || MODE == 2 || MODE == 4) { double w[1] = {v[0]}; block_reduce_publish<1>(w, partials, results, counter); }
    if (MODE == 3) block_reduce_publish<2>(v, partials, results, counter);
}

// band-indexed copies of the E arrays (update head of the fused iteration)
__global__ void kf_band_index(Grid g, FoldDev fd, long long *Bq, int *Bidx, double *Bblk, int nBp)
{
    const int NK = 1 + 2 * g.N;
    for (int bo = blockIdx.x * blockDim.x + threadIdx.x; bo < fd.nB; bo += gridDim.x * blockDim.x) {
        const int e = fd.EofB[bo];
        Bidx[bo] = e;
        if (e < 0) { Bq[bo] = 0; continue; }      // band cell of a ghost plane: no row on this rank
        Bq[bo] = fd_q(fd, fd.Ecell[e]);
        for (int kk = 0; kk < 2 * g.N; ++kk) {
            Bidx[(size_t)(1 + kk) * nBp + bo] = fd.EnbrB[(size_t)kk * fd.nEp + e];
            Bidx[(size_t)(1 + 2 * g.N + kk) * nBp + bo] = fd.EnbrE[(size_t)kk * fd.nEp + e];
        }
        for (int j = 0; j < NK * 9; ++j) Bblk[(size_t)j * nBp + bo] = fd.Eblk[(size_t)j * fd.nEp + e];
    }
}

// ---- band preconditioner -------------------------------------------------------------------------------------------------------
// The spectrum of M^ is the bulk interval [1/(1+2N theta dt/h^2 ...)] plus a few low modes localised on neighbouring cut cells (their w
// unknowns are coupled across faces; tests/experiments/krylov_experiment3.py).  They are removed by a low-degree Chebyshev polynomial of the band
// block M^_BB (all unknowns of the band cells) used as preconditioner on the band only: z = r outside the band, z_B = q(M^_BB) r_B.
// out[c][bo] = ca x_c + cb (M^_BB x)_c for every OWNED band cell; publishes sum_c x_c out_c  (x read from an FVec)
// pfix (fused iteration): before `out` (= dz) is overwritten, the OLD correction is added to the bulk entries of pfix on the band cells of the fused
// kernel's tiles (p_k = p' + dz, see kf_apply_band).
template <int N, int LPC>
__global__ void __launch_bounds__(256) kf_band_poly(Grid g, FoldDev fd, FVec x, double *out, double ca, double cb, double *partials, double *results,
                                                    unsigned *counter, const double *res, StopCrit stop, FVec pfix = FVec{{nullptr, nullptr, nullptr}})
{
    if (stop.sl_rr >= 0 && fold_done(res, stop)) return;
    double v[1] = {0.0};
    BAND_LOOP(LPC) {
        if (LPC == 1 && fd.EB[e] < 0) continue;                // fringe cells carry no preconditioner row
        double a0, a1, a2, x0, x1, xw;
        long long l; int bo;
        band_rows<N, true, LPC>(g, fd, e, x, a0, a1, a2, x0, x1, xw, l, bo, sub);
        if (sub != 0 || !live || bo < 0) continue;
        if (pfix.f[0] != nullptr && bo >= fd.nBlo && bo < fd.nBlo + fd.nBown && (fd.Efix[e] & 1u)) {
            const long long lq = fd_q(fd, l);
            pfix.f[0][lq] += out[(size_t)0 * fd.nB + bo];
            if (fd.nbulk > 1) pfix.f[1][lq] += out[(size_t)1 * fd.nB + bo];
        }
        const double o0 = ca * x0 + cb * a0, o1 = ca * x1 + cb * a1, o2 = ca * xw + cb * a2;
        out[(size_t)0 * fd.nB + bo] = o0;
        out[(size_t)1 * fd.nB + bo] = o1;
        out[(size_t)2 * fd.nB + bo] = o2;
        v[0] += x0 * o0 + x1 * o1 + xw * o2;
    }
    block_reduce_publish<1>(v, partials, results, counter);
}
// x_B = scale * in_B (band cells of an FVec), or x_B += in_B when add != 0
__global__ void kf_band_put(FoldDev fd, FVec x, const double *in, double scale, int add, const double *res, StopCrit stop)
{
    if (stop.sl_rr >= 0 && fold_done(res, stop)) return;
    for (int k = fd.nBlo + blockIdx.x * blockDim.x + threadIdx.x; k < fd.nBlo + fd.nBown; k += gridDim.x * blockDim.x) {
        const long long l = fd.Bcell[k];
        const double v0 = scale * in[(size_t)0 * fd.nB + k], v1 = scale * in[(size_t)1 * fd.nB + k], v2 = scale * in[(size_t)2 * fd.nB + k];
        const long long lq = fd_q(fd, l);
        if (add) { x.f[0][lq] += v0; if (fd.nbulk > 1) x.f[1][lq] += v1; x.f[2][k] += v2; }
        else { x.f[0][lq] = v0; if (fd.nbulk > 1) x.f[1][lq] = v1; x.f[2][k] = v2; }
    }
}
// deterministic pseudo-random start vector on the band (power iteration)
__global__ void kf_band_seed(FoldDev fd, double *out)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < fd.nB; k += gridDim.x * blockDim.x) {
        const long long l = fd.Bcell[k];
        for (int c = 0; c < 3; ++c) {
            unsigned long long h = (unsigned long long)(l * 3 + c) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
            out[(size_t)c * fd.nB + k] = 0.5 + (double)(h >> 11) * (1.0 / 9007199254740992.0);
        }
    }
}

// ---- vector kernels over the item list (dense active chunks + compact w) ------------------------------------------------------
#define FV_LOOP(I)                                                  \
    for (int it__ = blockIdx.x; it__ < (I).n; it__ += gridDim.x)    \
        if (const TileRec R__ = (I).rec[it__]; true)                \
            if (const int f = R__.f; true)                          \
                _Pragma("unroll") for (int k__ = 0; k__ < FU; ++k__) \
                    if (long long i = 0, q = 0; tile_cell((I), R__, k__, i, q))

// (FV_LOOP: i = index in reference-pitch arrays (capacities, masks, states, MVec), q = index in the re-pitched Krylov vectors (FVec))
__global__ void __launch_bounds__(FCH) kf_zero(Items I, FVec a) { FV_LOOP(I) { (void)i; a.f[f][q] = 0.0; } }
// deterministic pseudo-random start vector in (-1, 1) (power iteration for the top of the spectrum of M^)
__global__ void __launch_bounds__(FCH) kf_seed(Items I, FVec a)
{
    FV_LOOP(I) {
        unsigned long long h = (unsigned long long)(i * 3 + f + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
        a.f[f][q] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
}
__global__ void __launch_bounds__(FCH) kf_copy2(Items I, FVec a, FVec b, FVec c) { FV_LOOP(I) { (void)i; const double v = a.f[f][q]; b.f[f][q] = v; c.f[f][q] = v; } }
// r = b - q ; publishes (r, r)
// first residual of a solve, one pass: r = b - q (q == nullptr fields: r = b), copies of r into p (and r0), publishes ((b, b), (r, r))
// pzero: p = 0 instead of r (the fused CG iteration forms its first direction itself: p_0 = z_0 + beta_0 * 0)
__global__ void __launch_bounds__(FCH) kf_resid(Items I, FVec b, FVec qv, int have_q, FVec r, FVec p, FVec r0, int have_r0, int pzero, double *partials, double *results,
                                                unsigned *counter)
{
    double v[2] = {0.0, 0.0};
    FV_LOOP(I) {
        (void)i;
        const double bv = b.f[f][q];
        const double x = have_q ? bv - qv.f[f][q] : bv;
        r.f[f][q] = x;
        p.f[f][q] = pzero ? 0.0 : x;
        if (have_r0) r0.f[f][q] = x;
        v[0] += bv * bv;
        v[1] += x * x;
    }
    block_reduce_publish<2>(v, partials, results, counter);   // results = FS_BB, FS_RR0 (adjacent)
}
__global__ void __launch_bounds__(FCH) kf_dot(Items I, FVec a, FVec b, double *partials, double *results, unsigned *counter)
{
    double v[1] = {0.0};
    FV_LOOP(I) { (void)i; v[0] += a.f[f][q] * b.f[f][q]; }
    block_reduce_publish<1>(v, partials, results, counter);
}
// CG: r -= alpha q ; publishes (rho_new, rr) = ((r, r), (r, r)).  (x += alpha p is done by kf_cg_p, which reads p anyway.)
__global__ void __launch_bounds__(FCH) kf_cg_update(Items I, double *res, int sl_rho, int sl_new, FVec q, FVec r, double *partials, unsigned *counter, StopCrit stop)
{
    if (fold_done(res, stop)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) res[FS_ITERS] += 1.0;
    const double alpha = safe_div(rho_at(res, sl_rho), res[FS_SIG_D] + res[FS_SIG_B]);
    double v[2] = {0.0, 0.0};
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        const int f = R.f;
        const double *__restrict__ qf = f == 0 ? q.f[0] : (f == 1 ? q.f[1] : q.f[2]);
        double *__restrict__ rf = f == 0 ? r.f[0] : (f == 1 ? r.f[1] : r.f[2]);
        long long i[FU], iq[FU]; bool ok[FU]; double qv[FU], rv[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, i[k], iq[k]);
            if (ok[k]) { qv[k] = qf[iq[k]]; rv[k] = rf[iq[k]]; }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k)
            if (ok[k]) {
                const double rn = rv[k] - alpha * qv[k];
                rf[iq[k]] = rn;
                v[0] += rn * rn;
            }
    }
    v[1] = v[0];
    block_reduce_publish<2>(v, partials, res + sl_new, counter);
}
// p = z + beta p with z = r + dz on the band cells (dz = (q(M^_BB) - 1) r_B from kf_band_poly; nullptr: no band preconditioner).
// Also carries the (rho, rr) pair forward when the iteration was skipped by the stopping test (stop_old).
__global__ void __launch_bounds__(FCH) kf_cg_p(Items I, double *res, int sl_rho, int sl_new, FVec r, FVec p, FVec x, const double *__restrict__ dz,
                                               const int *__restrict__ bord, int nB, StopCrit stop_old, StopCrit stop)
{
    if (fold_done(res, stop_old)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { res[sl_new] = res[sl_rho]; res[sl_new + 1] = res[sl_rho + 1]; res[sl_new + 2] = res[sl_rho + 2]; res[sl_new + 3] = res[sl_rho + 3]; }
        return;
    }
    // x += alpha p_old belongs to this iteration even when it is the one that converged; the new direction is only needed otherwise
    const bool last = fold_done(res, stop);
    const double alpha = safe_div(rho_at(res, sl_rho), res[FS_SIG_D] + res[FS_SIG_B]);
    const double beta = safe_div(rho_at(res, sl_new), rho_at(res, sl_rho));
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        const int f = R.f;
        const double *__restrict__ rf = f == 0 ? r.f[0] : (f == 1 ? r.f[1] : r.f[2]);
        double *__restrict__ pf = f == 0 ? p.f[0] : (f == 1 ? p.f[1] : p.f[2]);
        double *__restrict__ xf = f == 0 ? x.f[0] : (f == 1 ? x.f[1] : x.f[2]);
        const bool band_tile = dz != nullptr && (f == 2 || (I.uni[it] & 2));
        long long i[FU], iq[FU]; bool ok[FU]; double pv[FU], rv[FU], xv[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            ok[k] = tile_cell(I, R, k, i[k], iq[k]);
            if (ok[k]) { pv[k] = pf[iq[k]]; xv[k] = xf[iq[k]]; rv[k] = last ? 0.0 : rf[iq[k]]; }
        }
        if (band_tile && !last) {
#pragma unroll
            for (int k = 0; k < FU; ++k)
                if (ok[k]) {
                    const long long bo = f == 2 ? i[k] : (long long)bord[i[k]];
                    if (bo >= 0) rv[k] += dz[(size_t)f * nB + bo];
                }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k)
            if (ok[k]) {
                xf[iq[k]] = xv[k] + alpha * pv[k];
                if (!last) pf[iq[k]] = rv[k] + beta * pv[k];
            }
    }
}
// BiCGSTAB
__global__ void __launch_bounds__(FCH) kf_bicg_s(Items I, const double *res, int sl_rho, FVec r, FVec v, FVec s, StopCrit stop)
{
    if (fold_done(res, stop)) return;
    const double alpha = safe_div(res[sl_rho], res[FS_SIG_D] + res[FS_SIG_B]);
    FV_LOOP(I) { (void)i; s.f[f][q] = r.f[f][q] - alpha * v.f[f][q]; }
}
__global__ void __launch_bounds__(FCH) kf_bicg_xr(Items I, double *res, int sl_rho, int sl_new, FVec p, FVec s, FVec t, FVec r0, FVec x, FVec r, double *partials,
                                                  unsigned *counter, StopCrit stop)
{
    if (fold_done(res, stop)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) res[FS_ITERS] += 1.0;
    const double alpha = safe_div(res[sl_rho], res[FS_SIG_D] + res[FS_SIG_B]);
    const double omega = safe_div(res[FS_TS_D] + res[FS_TS_B], res[FS_TT_D] + res[FS_TT_B]);
    double v[2] = {0.0, 0.0};
    FV_LOOP(I) {
        (void)i;
        const double sv = s.f[f][q];
        x.f[f][q] += alpha * p.f[f][q] + omega * sv;
        const double rn = sv - omega * t.f[f][q];
        r.f[f][q] = rn;
        v[0] += r0.f[f][q] * rn;
        v[1] += rn * rn;
    }
    block_reduce_publish<2>(v, partials, res + sl_new, counter);
}
__global__ void __launch_bounds__(FCH) kf_bicg_p(Items I, const double *res, int sl_rho, int sl_new, FVec r, FVec v, FVec p, StopCrit stop)
{
    if (fold_done(res, stop)) return;
    const double alpha = safe_div(res[sl_rho], res[FS_SIG_D] + res[FS_SIG_B]);
    const double omega = safe_div(res[FS_TS_D] + res[FS_TS_B], res[FS_TT_D] + res[FS_TT_B]);
    const double beta = safe_div(res[sl_new], res[sl_rho]) * safe_div(alpha, omega);
    FV_LOOP(I) { (void)i; p.f[f][q] = r.f[f][q] + beta * (p.f[f][q] - omega * v.f[f][q]); }
}

// iteration skipped by the stopping test: copy the (rho, rr) pair forward
__global__ void kf_carry_pair(double *res, int sl_old, int sl_new, StopCrit stop)
{
    if (threadIdx.x == 0 && fold_done(res, stop)) {
        res[sl_new] = res[sl_old]; res[sl_new + 1] = res[sl_old + 1];
        res[sl_new + 2] = res[sl_old + 2];
    }
}

// ---- transforms between the reference's unknowns / rows and the scaled ones ----------------------------------------------------
__device__ __forceinline__ double fold_rowscale(const FoldDev &fd, int p, long long l) { return fd.s[p] / (fd.c * D_at(fd.ph[p], l)); }

// b^ = L^-1 (rowscale . b): dense part (band cells get 0 here, the band kernel overwrites them)
__global__ void __launch_bounds__(FCH) kf_to_scaled_dense(FoldDev fd, Items I, MVec b, FVec bh)
{
    FV_LOOP(I) { if (f < 2) bh.f[f][q] = fd.sc[f][i] * fold_rowscale(fd, f, i) * b.f[f][i]; }
}
__global__ void kf_to_scaled_band(FoldDev fd, MVec b, FVec bh)
{
    for (int k = fd.nBlo + blockIdx.x * blockDim.x + threadIdx.x; k < fd.nBlo + fd.nBown; k += gridDim.x * blockDim.x) {
        const long long l = fd.Bcell[k];
        double I5[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) I5[q] = fd.Linv[(size_t)q * fd.nB + k];
        const double v0 = I5[0] != 0.0 ? fold_rowscale(fd, 0, l) * b.f[0][l] : 0.0;
        const double v1 = (fd.nbulk > 1 && I5[1] != 0.0) ? fold_rowscale(fd, 1, l) * b.f[1][l] : 0.0;
        const double vw = fd.wrow * b.f[fd.nbulk][l];
        const long long lq = fd_q(fd, l);
        bh.f[0][lq] = I5[0] * v0;
        if (fd.nbulk > 1) bh.f[1][lq] = I5[1] * v1;
        bh.f[2][k] = I5[2] * v0 + I5[3] * v1 + I5[4] * vw;
    }
}
// x^ = L^T x0 with x0 the extrapolated initial guess sum_j c_j T^(n-j) (GuessSpec, assemble.cuh), evaluated on the fly
__device__ __forceinline__ double guess_at(const GuessSpec &gs, long long l)
{
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < PB_MAXHIST; ++j)
        if (j < gs.m) v += gs.c[j] * gs.T[j][l];
    return v;
}
__global__ void __launch_bounds__(FCH) kf_guess_dense(FoldDev fd, Items I, GuessSpec g0, GuessSpec g1, FVec xh)
{
    FV_LOOP(I) {
        if (f < 2) {
            const double s = fd.sc[f][i];     // non-zero exactly on the free, non-band unknowns
            xh.f[f][q] = s != 0.0 ? guess_at(f == 0 ? g0 : g1, i) / s : 0.0;
        }
    }
}
// Fused prologue of a step without an explicit operator part (BE, steady) on the folded path: one pass over the active tiles gives
//   b   = cV V T^n + V (wf0 f(t_n) + wf1 f(t_n+1))   the bulk rows of b_*_unstead_diff / b_*_stead_diff (src/solver/diffusion.jl:45-58,146-161,
//                                                     243-265,391-420) without their known part (k_rhs_known_* adds it on its row list),
//   b^  = sc rowscale b                               (kf_to_scaled_dense),
//   x^0 = guess / sc                                  (kf_guess_dense)
// instead of three kernels over all cells of every phase (k_rhs_*, kf_to_scaled_dense, kf_guess_dense).
struct RhsSrc { SrcSpec f0[2], f1[2]; const double *Tw[2]; const unsigned char *m[2]; };
// Crank-Nicolson (vexp != null): the explicit operator part of a row that couples to no eliminated value is the folded operator applied
// to the old state, V T - c (D G'W!(G T + H Tg)) = 2 V T - (A T)_i with (L^-1 S A T)_i = (M^ x^n)_i, x^n = L^T T^n on the free unknowns:
// b^ = sc rowscale (2 V T^n + V (wf0 f^n + wf1 f^n+1)) - vexp, one folded apply instead of the unfolded stencil on every cell.  Rows that
// do couple to eliminated values (the k_mark_known_rows list: band, border- and interface-adjacent rows) are evaluated unfolded by
// k_rhs_* in list mode afterwards and refreshed by kf_to_scaled_list / kf_to_scaled_band; b is meaningful on those rows only.
__global__ void __launch_bounds__(FCH, 4) kf_rhs_dense(FoldDev fd, Items I, StepCoef sc, RhsSrc S, GuessSpec g0, GuessSpec g1, MVec b, FVec bh, FVec xh, FVec vexp)
{
    for (int it = blockIdx.x; it < I.n; it += gridDim.x) {
        const TileRec R = I.rec[it];
        if (R.f >= 2) continue;
        // the field is one per tile: every per-field pointer is selected here (no dynamically indexed kernel parameters)
        const bool f1 = R.f == 1;
        const unsigned char *__restrict__ m = f1 ? S.m[1] : S.m[0];
        const double *__restrict__ Vf = f1 ? fd.ph[1].V : fd.ph[0].V;
        const double *__restrict__ Da = f1 ? fd.ph[1].Darr : fd.ph[0].Darr;
        const double *__restrict__ Tw = f1 ? S.Tw[1] : S.Tw[0];
        const double *__restrict__ scf = f1 ? fd.sc[1] : fd.sc[0];
        const double *__restrict__ fa0 = f1 ? S.f0[1].arr : S.f0[0].arr, *__restrict__ fa1 = f1 ? S.f1[1].arr : S.f1[0].arr;
        const double fc0 = f1 ? S.f0[1].cst : S.f0[0].cst, fc1 = f1 ? S.f1[1].cst : S.f1[0].cst;
        const double sF = f1 ? fd.s[1] : fd.s[0], Dc = f1 ? fd.ph[1].Dc : fd.ph[0].Dc;
        double *__restrict__ bo = f1 ? b.f[1] : b.f[0], *__restrict__ bho = f1 ? bh.f[1] : bh.f[0], *__restrict__ xo = f1 ? xh.f[1] : xh.f[0];
        const double *__restrict__ ve = f1 ? vexp.f[1] : vexp.f[0];
        const double cT = ve ? 2.0 * sc.cV : sc.cV;
        long long idx[FU], idq[FU];
        bool ok[FU];
        double s[FU], v[FU], gs[FU], ex[FU];
#pragma unroll
        for (int k = 0; k < FU; ++k) ok[k] = tile_cell(I, R, k, idx[k], idq[k]);
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            s[k] = 0.0; v[k] = 0.0; gs[k] = 0.0; ex[k] = 0.0;
            if (ok[k]) {
                // every load of the cell is issued unconditionally (one round trip): the mask and the scale only select afterwards
                const long long i = idx[k];
                s[k] = scf[i];
                const bool fr = (m[i] & MB_FREE) != 0;
                const double V = Vf[i], T = Tw[i], q0 = fa0 ? fa0[i] : fc0, q1 = fa1 ? fa1[i] : fc1;
                if (ve) ex[k] = ve[idq[k]];
                double gsum = 0.0;
#pragma unroll
                for (int j = 0; j < PB_MAXHIST; ++j)
                    if (j < g0.m) gsum += (f1 ? g1.c[j] : g0.c[j]) * (f1 ? g1.T[j] : g0.T[j])[i];
                v[k] = fr ? cT * V * T + V * (sc.wf0 * q0 + sc.wf1 * q1) : 0.0;
                gs[k] = s[k] != 0.0 ? gsum : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < FU; ++k) {
            if (ok[k]) {
                const long long i = idx[k], iq = idq[k];
                bo[i] = v[k];
                bho[iq] = s[k] != 0.0 ? s[k] * (sF / (fd.c * (Da ? Da[i] : Dc))) * v[k] - ex[k] : 0.0;   // = sc fold_rowscale b (- M^ x^n)
                if (g0.m > 0) xo[iq] = s[k] != 0.0 ? gs[k] / s[k] : 0.0;
            }
        }
    }
}
// b^ of the listed rows again, after k_rhs_known_* has subtracted their known part from b (band rows: kf_to_scaled_band)
__global__ void kf_to_scaled_list(FoldDev fd, const long long *__restrict__ list, int n, MVec b, FVec bh)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const long long l = list[k];
        for (int f = 0; f < fd.nbulk; ++f) {
            const double s = fd.sc[f][l];
            if (s != 0.0) bh.f[f][fd_q(fd, l)] = s * fold_rowscale(fd, f, l) * b.f[f][l];
        }
    }
}
__global__ void kf_guess_band(FoldDev fd, GuessSpec g0, GuessSpec g1, GuessSpec gw, FVec xh)
{
    for (int k = fd.nBlo + blockIdx.x * blockDim.x + threadIdx.x; k < fd.nBlo + fd.nBown; k += gridDim.x * blockDim.x) {
        const long long l = fd.Bcell[k];
        double I5[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) I5[q] = fd.Linv[(size_t)q * fd.nB + k];
        const double l00 = I5[0] != 0.0 ? 1.0 / I5[0] : 0.0, l11 = I5[1] != 0.0 ? 1.0 / I5[1] : 0.0, l22 = 1.0 / I5[4];
        const double l20 = -I5[2] * l00 * l22, l21 = -I5[3] * l11 * l22;
        const double x0 = I5[0] != 0.0 ? guess_at(g0, l) : 0.0, x1 = (fd.nbulk > 1 && I5[1] != 0.0) ? guess_at(g1, l) : 0.0, xw = guess_at(gw, l);
        const long long lq = fd_q(fd, l);
        xh.f[0][lq] = l00 * x0 + l20 * xw;
        if (fd.nbulk > 1) xh.f[1][lq] = l11 * x1 + l21 * xw;
        xh.f[2][k] = l22 * xw;
    }
}
// x = L^-T x^
__global__ void __launch_bounds__(FCH) kf_from_scaled_dense(FoldDev fd, Items I, FVec xh, MVec x)
{
    FV_LOOP(I) { if (f < 2) x.f[f][i] = fd.sc[f][i] * xh.f[f][q]; }
}
__global__ void kf_from_scaled_band(FoldDev fd, FVec xh, MVec x)
{
    for (int k = fd.nBlo + blockIdx.x * blockDim.x + threadIdx.x; k < fd.nBlo + fd.nBown; k += gridDim.x * blockDim.x) {
        const long long l = fd.Bcell[k];
        double I5[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) I5[q] = fd.Linv[(size_t)q * fd.nB + k];
        const double hw = xh.f[2][k];
        const long long lq = fd_q(fd, l);
        x.f[0][l] = I5[0] * xh.f[0][lq] + I5[2] * hw;
        if (fd.nbulk > 1) x.f[1][l] = I5[1] * xh.f[1][lq] + I5[3] * hw;
        x.f[fd.nbulk][l] = I5[4] * hw;
    }
}

// =================================================================================================================================
// host side
// =================================================================================================================================
struct FoldGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0, applies = 0; };
struct FoldSys {
    bool built = false;
    FoldDev d;
    // owned device memory
    double *sc[2] = {}, *off[2][PB_MAXD] = {};
    long long *Bcell = nullptr, *Ecell = nullptr;
    int *bord = nullptr, *EB = nullptr, *EnbrB = nullptr, *items = nullptr;
    double *Linv = nullptr, *Eblk = nullptr;
    unsigned char *Efix = nullptr;
    int *eord = nullptr, *EofB = nullptr, *EnbrE = nullptr, *Bidx = nullptr;   // band heads of the fused iteration (fold2.cuh)
    long long *Bq = nullptr;
    double *Bblk = nullptr;
    double *ya = nullptr;
    int nitems = 0;
    unsigned char *uni = nullptr;
    double *ucoef = nullptr;
    TileRec *rec = nullptr;
    double *dz = nullptr;          // [3][nB] band preconditioner correction z_B - r_B
    bool prec = false;             // band preconditioner available
    double pa0 = 1.0, pa1 = 0.0;   // z_B = pa0 r_B + pa1 M^_BB r_B
    double band_lmin = 0.0, band_lmax = 0.0;
    std::map<int, FoldGraph> graphs;        // chunks of Krylov iterations captured as CUDA graphs, by chunk length (single GPU)
    double graph_key[5] = {};
    int last_iters = 0;                     // iteration count of the previous solve (sizes the first chunk of the next one)
    long long cells_uniform = 0, cells_general = 0;   // cells of tiles applied with constant / streamed coefficients (this rank)
    long long cells_fast = 0;                         // cells of full tiles with constant coefficients (the apply kernel's staged interior branch)
    Items I;                                // every item, index order (vector kernels)
    Items IA;                               // bulk tiles in cost-class order (operator apply)
    Items IAg;                              // the ghost class (box reaches a neighbour rank's ghost plane), every kind of tile
    Items IAf, IAgen;                       // interior class: tiles of the pipelined kernel (all cells valid, constant coefficients) / the rest
    Items IFall, IGall;                     // the same split over both classes (plain applies with the halo already exchanged)
    Items IG1;                              // ghost-class tiles + compact interface unknowns: pointwise p / x update of the fused iteration
    Items IG1nw;                            // ghost-class tiles only (band heads update the interface unknowns)
    bool bandfuse_ok = false;               // band heads usable (collective: the same on every rank)
    std::vector<void *> list_mem;           // device arrays behind the sub-lists
    FVec x, b, r, p, v, r0, s, t, z;
    FVec p2 = {}, zz = {};                  // fused iteration: second search-direction buffer, preconditioned residual (polynomial)
    FVec r2 = {};                           // band heads: r2.f[2] is the second buffer of the interface part of r; r2.f[0], r2.f[1] alias r (never freed through r2)
    double *rE = nullptr;                   // band heads: [2 parities][2][nEp] compact copy of r on the E cells
    bool have_p2 = false, have_zz = false, have_r2 = false;
    bool tma_ok = false;                    // the Krylov vectors are describable to TMA (fold2.cuh)
    bool pipe = false;                      // ... and the interior constant-coefficient tiles go through the pipelined kernel (kf3_apply)
    Items IAi_all;                          // interior class, every kind of tile (fused iteration without the pipelined kernel)
    std::map<std::pair<const double *, int>, CUtensorMap> tmaps;
    bool have_bicg = false, have_z = false;
    int poly_m = 0;                         // degree of the polynomial preconditioner q(M^) (0: none)
    double poly_lo = 0.0, poly_hi = 0.0;    // Chebyshev interval of the bulk spectrum
    long long wcap = 0;
    long long P0 = 0, nlocq = 0, planeq = 0;   // x pitch, size and slab-plane size of the re-pitched Krylov vectors
    double key[8] = {};   // coefficient set the system was built for
};

static void fold_free_vec(FVec &a) { for (int f = 0; f < 3; ++f) { if (a.f[f]) cudaFree(a.f[f]); a.f[f] = nullptr; } }
static void fold_free(FoldSys &F)
{
    for (auto &kv : F.graphs) cudaGraphExecDestroy(kv.second.exec);
    F.graphs.clear();
    memset(F.graph_key, 0, sizeof(F.graph_key));
    for (int p = 0; p < 2; ++p) { dev_free(F.sc[p]); for (int d = 0; d < PB_MAXD; ++d) dev_free(F.off[p][d]); }
    if (F.Bcell) cudaFree(F.Bcell); if (F.Ecell) cudaFree(F.Ecell); if (F.bord) cudaFree(F.bord); if (F.EB) cudaFree(F.EB);
    if (F.uni) cudaFree(F.uni); if (F.ucoef) cudaFree(F.ucoef); if (F.rec) cudaFree(F.rec); if (F.dz) cudaFree(F.dz); F.dz = nullptr; F.prec = false; F.uni = nullptr; F.ucoef = nullptr; F.rec = nullptr;
    if (F.EnbrB) cudaFree(F.EnbrB); if (F.items) cudaFree(F.items); if (F.Linv) cudaFree(F.Linv); if (F.Eblk) cudaFree(F.Eblk);
    if (F.Efix) cudaFree(F.Efix); F.Efix = nullptr;
    if (F.eord) cudaFree(F.eord); if (F.EofB) cudaFree(F.EofB); if (F.ya) cudaFree(F.ya); if (F.EnbrE) cudaFree(F.EnbrE); F.eord = F.EofB = F.EnbrE = nullptr; F.ya = nullptr;
    if (F.Bidx) cudaFree(F.Bidx); if (F.Bq) cudaFree(F.Bq); if (F.Bblk) cudaFree(F.Bblk); F.Bidx = nullptr; F.Bq = nullptr; F.Bblk = nullptr;
    F.Bcell = F.Ecell = nullptr; F.bord = F.EB = F.EnbrB = F.items = nullptr; F.Linv = F.Eblk = nullptr;
    for (void *m : F.list_mem) cudaFree(m);
    F.list_mem.clear();
    F.tmaps.clear();
    F.r2.f[0] = F.r2.f[1] = nullptr;        // (aliases of r)
    if (F.rE) cudaFree(F.rE); F.rE = nullptr;
    FVec *vs[] = {&F.x, &F.b, &F.r, &F.p, &F.v, &F.r0, &F.s, &F.t, &F.z, &F.p2, &F.zz, &F.r2};
    for (FVec *a : vs) fold_free_vec(*a);
    F.built = false; F.have_bicg = false; F.have_z = false; F.have_p2 = false; F.have_zz = false; F.have_r2 = false; F.tma_ok = false;
}

static const int PB_NCCL_UINT8 = 1;
// ghost planes of byte fields (masks)
static int halo_exchange_bytes(pb200_ctx *ctx, const Grid &g, unsigned char *const *fields, int nf)
{
    if (ctx->nranks == 1) return PB200_OK;
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int f = 0; f < nf; ++f) {
        unsigned char *p = fields[f];
        if (!p) continue;
        size_t cnt = (size_t)g.plane;
        if (ctx->rank > 0) {
            NCCL_TRY(ctx, g_nccl.Send(p + g.plane, cnt, PB_NCCL_UINT8, ctx->rank - 1, ctx->comm, ctx->stream));
            NCCL_TRY(ctx, g_nccl.Recv(p, cnt, PB_NCCL_UINT8, ctx->rank - 1, ctx->comm, ctx->stream));
        }
        if (ctx->rank < ctx->nranks - 1) {
            NCCL_TRY(ctx, g_nccl.Send(p + (long long)(g.lz - 2) * g.plane, cnt, PB_NCCL_UINT8, ctx->rank + 1, ctx->comm, ctx->stream));
            NCCL_TRY(ctx, g_nccl.Recv(p + (long long)(g.lz - 1) * g.plane, cnt, PB_NCCL_UINT8, ctx->rank + 1, ctx->comm, ctx->stream));
        }
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    return PB200_OK;
}

// halo of compact band arrays: `na` arrays of nB doubles (array a at base + a * nB).  The entries of the first / last OWNED plane
// are contiguous ranges (the list is sorted by cell index) and match the neighbour's ghost prefix / suffix entry for entry.
struct BandHalo { int lo_send0 = 0, lo_sendn = 0, hi_send0 = 0, hi_sendn = 0; };
static int fold_band_ranges(pb200_ctx *ctx, const Grid &g, const std::vector<long long> &hB, int nBlo, int nBown, BandHalo *bh)
{
    // first owned plane: cells [plane, 2 plane); last owned plane: [(lz-2) plane, (lz-1) plane)
    const long long p = g.plane;
    int i = nBlo;
    bh->lo_send0 = i;
    while (i < nBlo + nBown && hB[i] < 2 * p) ++i;
    bh->lo_sendn = i - bh->lo_send0;
    int j = nBlo + nBown;
    while (j > nBlo && hB[j - 1] >= (long long)(g.lz - 2) * p) --j;
    bh->hi_send0 = j;
    bh->hi_sendn = nBlo + nBown - j;
    (void)ctx;
    return PB200_OK;
}
static int fold_band_halo(pb200_ctx *ctx, const FoldSys &F, const BandHalo &bh, double *base, int na)
{
    if (ctx->nranks == 1 || F.d.nB == 0) return PB200_OK;
    const int nB = F.d.nB, nBlo = F.d.nBlo, nBown = F.d.nBown, nBhi = nB - nBlo - nBown;
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int a = 0; a < na; ++a) {
        double *p = base + (size_t)a * nB;
        if (ctx->rank > 0) {
            if (bh.lo_sendn) NCCL_TRY(ctx, g_nccl.Send(p + bh.lo_send0, (size_t)bh.lo_sendn, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, ctx->stream));
            if (nBlo) NCCL_TRY(ctx, g_nccl.Recv(p, (size_t)nBlo, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, ctx->stream));
        }
        if (ctx->rank < ctx->nranks - 1) {
            if (bh.hi_sendn) NCCL_TRY(ctx, g_nccl.Send(p + bh.hi_send0, (size_t)bh.hi_sendn, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, ctx->stream));
            if (nBhi) NCCL_TRY(ctx, g_nccl.Recv(p + nBlo + nBown, (size_t)nBhi, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, ctx->stream));
        }
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    return PB200_OK;
}
