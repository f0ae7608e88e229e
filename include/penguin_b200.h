/*
 * penguin_b200.h -- C ABI of libpenguin_b200.so
 *
 * Drop-in boundary for the unsteady cut-cell diffusion hot path of Fastaxx/Penguin.jl.  The reference has no
 * FFI of its own (pure Julia): the boundary is its exported Julia API.  Each entry point
 * below is what a Julia shim `ccall`s to replace one reference method (cited per function as file:line under
 * /root/reference).  INTEGRATION.md shows the Julia-side binding.
 *
 * Conventions
 *  - every per-cell array has the reference's padded length n = prod(n_i + 1), x fastest
 *    (src/capacity.jl:167-175, src/solver.jl:362-372); all values are IEEE double;
 *  - host buffers are owned by the caller and never retained; device memory lives behind the opaque handles;
 *  - every function returns 0 on success, a PB200_E* code otherwise; pb200_last_error() gives the message;
 *  - there is NO CPU fallback: without a CUDA device pb200_init fails with PB200_ENODEV;
 *  - multi-GPU: one process per GPU (pb200_init_dist), the slowest grid dimension is slab-partitioned over the
 *    ranks; per-cell host arrays then hold the rank's OWNED planes only (pb200_grid_local tells which).
 */
#ifndef PENGUIN_B200_H
#define PENGUIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_OK 0
#define PB200_EINVAL 1   /* bad argument                                   */
#define PB200_ENODEV 2   /* no CUDA device / driver                        */
#define PB200_ECUDA 3    /* CUDA runtime error                             */
#define PB200_ENCCL 4    /* NCCL error or libnccl not loadable             */
#define PB200_ENOTCONV 5 /* Krylov solve hit maxit (state still updated)   */
#define PB200_EUNSUPPORTED 6

typedef struct pb200_ctx pb200_ctx;
typedef struct pb200_capacity pb200_capacity;
typedef struct pb200_ops pb200_ops;
typedef struct pb200_solver pb200_solver;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
int pb200_init(pb200_ctx **ctx, int device);
/* one process per GPU; `nccl_id` = 128 bytes from pb200_nccl_unique_id() of rank 0, broadcast by the host. */
int pb200_nccl_unique_id(char id[128]);
int pb200_init_dist(pb200_ctx **ctx, int device, int rank, int nranks, const char nccl_id[128]);
/* ONE process driving ndev GPUs of one box (the reference's scripts are single-process Julia): the returned handle stands for the whole team.
 * Every call below then takes and returns per-cell host arrays of the reference's GLOBAL padded length -- the library cuts them into slabs of the
 * slowest dimension, runs one host thread per GPU (NCCL / peer-memory exchanges between them) and reassembles the results.  An unchanged
 * single-process script reaches N > 1 GPUs by calling this instead of pb200_init.  (pb200_solver_get_state_async is synchronous on a team handle.) */
int pb200_init_multi(pb200_ctx **ctx, const int *devices, int ndev);
int pb200_finalize(pb200_ctx *ctx);
const char *pb200_last_error(pb200_ctx *ctx); /* ctx may be NULL: last error of the calling thread */
int pb200_sync(pb200_ctx *ctx);
/* number of kernels this library has launched on ctx since init (bench.py reports it as gpu_launches) */
int64_t pb200_launch_count(pb200_ctx *ctx);
/* the CUDA stream every kernel of ctx is launched on (cudaStream_t as an integer), for event timing by the host */
uint64_t pb200_stream(pb200_ctx *ctx);
/* enable != 0: bracket every operator-apply launch of pb200_solver_step with CUDA events on the launching stream and
 * report their summed device time in pb200_step_stats.apply_ms (the roofline measurement of bench.py) */
int pb200_set_profiling(pb200_ctx *ctx, int enable);

/* ---- level-set descriptors (GPU-evaluable bodies; a Julia closure cannot run on the device) ---------------- */
#define PB200_LS_BALLS 0     /* phi = min_k |x - c_k| - r_k, disjoint balls (interval / circle / sphere) */
#define PB200_LS_HALFSPACE 1 /* phi = x[dim] - c                                                         */
typedef struct {
    int kind;
    int nballs;
    const double *centers; /* nballs * ndim, ball-major   */
    const double *radii;   /* nballs                      */
    int fluid_inside;      /* 1: fluid = {phi < 0}; 0: sign flip (the reference's `-(...)` bodies) */
    int hs_dim;
    double hs_c;
} pb200_levelset;

/* ---- Capacity: replaces Capacity(body, mesh; method="VOFI") -- src/capacity.jl:51-64, 81-123, 137-197 -------- */
/* mesh = Penguin.Mesh(n, L, x0) (src/mesh.jl:41-79); geometry cells are [nodes_j, nodes_j+1], nodes = x0+(j+1/2)h */
int pb200_capacity_create(pb200_ctx *ctx, int ndim, const int *n, const double *x0, const double *L,
                          const pb200_levelset *ls, int compute_centroids, pb200_capacity **cap);
/* arrays computed elsewhere (generic closures evaluated by the reference, parity staging): any pointer may be NULL
 * (treated as zeros). A, B, W, C_omega, C_gamma are component-major: ndim blocks of nloc doubles.            */
int pb200_capacity_import(pb200_ctx *ctx, int ndim, const int *n, const double *x0, const double *L,
                          const double *V, const double *Gamma, const double *cell_types, const double *A,
                          const double *B, const double *W, const double *C_omega, const double *C_gamma,
                          pb200_capacity **cap);
/* fills caller buffers (NULL = skip) with the fields of the Capacity struct, src/capacity.jl:25-36 */
int pb200_capacity_export(pb200_capacity *cap, double *V, double *Gamma, double *cell_types, double *A, double *B,
                          double *W, double *C_omega, double *C_gamma);
/* owned slab of this rank: padded plane range [k0, k1) of the slowest dimension and nloc = owned cells */
int pb200_capacity_local(pb200_capacity *cap, int *k0, int *k1, int64_t *nloc);
int pb200_capacity_destroy(pb200_capacity *cap);

/* ---- DiffusionOps: replaces DiffusionOps(cap) -- src/operators.jl:127-178 (matrix-free; G, H never assembled) -- */
int pb200_ops_create(pb200_capacity *cap, pb200_ops **ops);
/* grad(op, p) = W!(G p_omega + H p_gamma) -- src/operators.jl:20-23.  p: 2*nloc, out: ndim*nloc            */
int pb200_ops_grad(pb200_ops *ops, const double *p, double *out);
/* div(op, q_omega, q_gamma) = -(G'+H') q_omega + H' q_gamma -- src/operators.jl:30-34.  q_*: ndim*nloc   */
int pb200_ops_div(pb200_ops *ops, const double *q_omega, const double *q_gamma, double *out);
/* ConvectionOps(cap, u_omega, u_gamma) -- src/operators.jl:194-209: adds the advective operators C_d = D_p diag(S_m A_d u_omega_d) S_m and
 * K_d = diag(S_p H' u_gamma) to an operator handle (coefficient arrays on the device, applied matrix-free).  u_omega, u_gamma: ndim*nloc each.
 * A solver created on such operators is the reference's AdvectionDiffusion{Steady,Unsteady}{Mono,Diph} (src/solver/advectiondiffusion.jl:12-418):
 * bulk rows gain (sum_d C_d + 0.5 sum_d K_d) T_omega + 0.5 sum_d K_d T_gamma; solved with BiCGSTAB on the reference's rows (one GPU).          */
int pb200_ops_set_convection(pb200_ops *ops, const double *u_omega, const double *u_gamma);
/* the coefficient arrays behind it: cf (ndim*nloc: S_m A_d u_omega_d, so that C_d = D_p diag(cf_d) S_m) and kd (nloc: diag of 0.5 sum_d K_d); NULL = skip */
int pb200_ops_export_convection(pb200_ops *ops, double *cf, double *kd);
/* W! diagonal (1/W, 1.0 where W == 0 -- src/operators.jl:145-152), component-major ndim*nloc              */
int pb200_ops_export_wdag(pb200_ops *ops, double *wdag);
int pb200_ops_destroy(pb200_ops *ops);

/* ---- Solver: replaces Diffusion{Steady,Unsteady}{Mono,Diph} + solve_...! -- src/solver/diffusion.jl ---------- */
#define PB200_MONO 0
#define PB200_DIPH 1
#define PB200_STEADY 0
#define PB200_UNSTEADY 1
#define PB200_BC_NONE 0
#define PB200_BC_DIRICHLET 1
#define PB200_BC_NEUMANN 2
#define PB200_BC_ROBIN 3
#define PB200_BC_PERIODIC 4
/* border sides in the reference's key order (src/solver.jl:379-409):
 * left=(dim2,lo) right=(dim2,hi) bottom=(dim1,lo) top=(dim1,hi) backward=(dim3,lo) forward=(dim3,hi)         */
#define PB200_LEFT 0
#define PB200_RIGHT 1
#define PB200_BOTTOM 2
#define PB200_TOP 3
#define PB200_BACKWARD 4
#define PB200_FORWARD 5

typedef struct {
    int phase_type; /* PB200_MONO | PB200_DIPH          */
    int time_type;  /* PB200_STEADY | PB200_UNSTEADY    */
    pb200_ops *ops1, *ops2;
    /* diffusion coefficient D(C_omega): constant, or host array of nloc values (build_I_D, src/solver.jl:255-266) */
    double D1, D2;
    const double *D1_arr, *D2_arr;
    /* mono interface condition (build_I_bc, src/solver.jl:203-223): Dirichlet (1,0) Neumann (0,1) Robin (a,b) */
    int ifc_kind;
    double alpha, beta;
    /* diph jumps (src/boundary.jl:97-118): alpha1 Tg1 - alpha2 Tg2 = g ; beta1 q1 + beta2 q2 = Gamma2 h */
    double alpha1, alpha2, beta1, beta2;
} pb200_solver_desc;

int pb200_solver_create(pb200_ctx *ctx, const pb200_solver_desc *desc, pb200_solver **s);
/* border condition of one side (BC_border_mono!/diph!, src/solver.jl:417-580).  `values` (host, may be NULL =>
 * `value` everywhere) holds one entry per real cell of the side, the other dims x fastest.  May be called again
 * before any step (time-dependent border data).  Dirichlet: x = value.  Periodic: supported when the row pins a value (both sides
 * Periodic, or Periodic + Dirichlet).  Neumann: a real row in 1-D only (`value` = g), a no-op in >= 2-D like Robin (src/solver.jl:471-498). */
int pb200_solver_set_border(pb200_solver *s, int side, int kind, double value, const double *values);
/* state vector [T_omega; T_gamma] (2 nloc) or [T_omega1; T_gamma1; T_omega2; T_gamma2] (4 nloc) */
int pb200_solver_set_state(pb200_solver *s, const double *x);
int pb200_solver_get_state(pb200_solver *s, double *x);
/* Streaming variant for time loops that keep every state on the host (solver.states, src/solver/diffusion.jl:296,450): the copy of the
 * CURRENT state into `x` (pinned memory recommended) is queued on a separate copy stream and overlaps the next pb200_solver_step, whose
 * write-back of the new state waits for the copy to finish.  `x` holds the state after pb200_solver_wait_state (or the next _async call). */
int pb200_solver_get_state_async(pb200_solver *s, double *x);
int pb200_solver_wait_state(pb200_solver *s);

#define PB200_BE 0
#define PB200_CN 1
typedef struct {
    int scheme; /* PB200_BE | PB200_CN (ignored for steady) */
    double dt;
    /* bulk source f at t_n and t_n + dt per phase (build_source, src/solver.jl:283-286): constant or nloc array */
    double f_const[2][2];
    const double *f_arr[2][2]; /* [phase][0: t_n, 1: t_n+dt] */
    /* interface data (build_g_g, src/solver.jl:309-323): mono g(t_n), g(t_n+dt); diph [0] = scalar-jump g, [1] = flux-jump h */
    double g_const[2];
    const double *g_arr[2];
} pb200_step_in;

/* kernel classes of pb200_step_stats.kernel_ms / kernel_launches */
#define PB200_K_APPLY 0      /* operator apply on the bulk tiles (+ fused search-direction update on the fused path) */
#define PB200_K_UPDATE 1     /* residual update with fused dot products                                               */
#define PB200_K_PUPD 2       /* solution / search-direction update (unfused path)                                     */
#define PB200_K_BAND_APPLY 3 /* interface-band part of the operator                                                   */
#define PB200_K_BAND_PREC 4  /* interface-band preconditioner                                                         */
#define PB200_K_EXCHANGE 5   /* halo exchange + scalar reductions between ranks                                       */
#define PB200_K_PROLOGUE 6   /* per-step right-hand side, scaling, initial guess, first residual                      */
#define PB200_K_EPILOGUE 7   /* back-transform and state write-back                                                   */
#define PB200_K_NCLASS 8

#define PB200_KRYLOV_AUTO 0 /* CG on the folded (symmetrised) path and for mono; BiCGSTAB for diph on the generic path */
#define PB200_KRYLOV_CG 1
#define PB200_KRYLOV_BICGSTAB 2
/* PB200_PATH_FOLDED: symmetrised, per-cell block-Jacobi-scaled stencil with unit diagonal (csrc/fold.cuh) -- the fast path, used
 * whenever the jump / Robin coefficients allow it.  PB200_PATH_GENERIC: the reference's rows applied matrix-free with point-Jacobi
 * preconditioning (csrc/operators.cuh) -- any coefficients, and the cross-check of the folded path in the tests.          */
#define PB200_PATH_AUTO 0
#define PB200_PATH_GENERIC 1
#define PB200_PATH_FOLDED 2
/* PB200_PRECOND_DEFAULT: what DESIGN.md section 4 describes (block-Jacobi scaling + interface-band / Chebyshev polynomial preconditioners).
 * PB200_PRECOND_MG: geometric multigrid V-cycle (csrc/mg.cuh: capacities rebuilt on the meshes n/2, n/4, ..., cell-aggregation transfers,
 * Chebyshev smoothers) for systems whose condition number grows like n^2 -- the steady problems of src/solver/diffusion.jl:14-72.  Monophasic,
 * Dirichlet interface, constant D, capacities built by pb200_capacity_create, one rank; PB200_EUNSUPPORTED otherwise.                         */
#define PB200_PRECOND_DEFAULT 0
#define PB200_PRECOND_MG 1
typedef struct {
    int method;
    double rtol; /* stop when ||r|| <= max(rtol ||b||, atol)  (IterativeSolvers convention) */
    double atol;
    int maxit;
    int warm_start; /* 0: zero initial guess like the reference; 1: previous state; m >= 2: polynomial extrapolation through the last m states (m <= 5) */
    int check_every; /* convergence is tested on the host every this many iterations (>= 1) */
    int path;        /* PB200_PATH_*: which implementation of the solve runs */
    int precond;     /* PB200_PRECOND_*: preconditioner of the CG on the folded path */
} pb200_krylov_opts;

typedef struct {
    int iters;
    int converged;
    double rnorm;    /* final ||r||_2 of the reduced system */
    double bnorm;    /* ||b||_2 of the reduced system       */
    double solve_ms; /* device time of the Krylov loop (CUDA events) */
    double setup_ms; /* device time of RHS / border assembly */
    int64_t dof_bulk; /* active bulk unknowns (all ranks) -- SURVEY 8(d) definition of DOF */
    int64_t dof_ifc;  /* active interface unknowns (all ranks) */
    int64_t launches; /* kernels launched by this call */
    double apply_ms;        /* summed device time of the operator-apply launches (0 unless profiling is enabled) */
    int64_t apply_launches; /* number of operator-apply launches of this call */
    /* folded path, this rank: cells of the tiles the apply kernel processes with per-tile constant coefficients (no coefficient
     * arrays read) and with streamed coefficient arrays -- the census behind bench.py's algorithmic-byte count */
    int64_t apply_cells_uniform, apply_cells_general;
    /* profiling enabled: summed device time / launch count per kernel class of the Krylov loop (PB200_K_* below) */
    double kernel_ms[8];
    int64_t kernel_launches[8];
    /* cells of the tiles that take the apply kernel's staged interior branch (every cell valid AND constant coefficients) -- the parity
     * tests assert this is > 0 so that the branch the benchmark runs is the branch they compare with the oracle */
    int64_t apply_cells_fast;
    /* interface band of this rank: band cells (cells with an active interface unknown) and rows of the band kernels (band + fringe cells) */
    int64_t band_cells, band_rows;
} pb200_step_stats;

/* one solve: builds b from the device-resident state (b_*_unstead_diff / b_*_stead_diff), applies the border
 * rows, solves the reduced system, stores the new state on the device (solve_system!, src/solver.jl:158-188:
 * removed DOFs are exactly 0).  opts / stats may be NULL.                                                      */
int pb200_solver_step(pb200_solver *s, const pb200_step_in *in, const pb200_krylov_opts *opts, pb200_step_stats *stats);
/* check_convergence (src/convergence.jl:59-93) without a host round trip of the state: volume-weighted L^p norms (p > 0, or INFINITY) of
 * u_ana - T_omega of bulk field `phase` (0, or 1 for the second phase of a diphasic solver), u_ana[n] evaluated by the caller at C_omega;
 * out = {all fluid cells (full + cut), full, cut, empty}, classes by capacity.cell_types.  relative != 0: relative_lp_norm.                */
int pb200_solver_error_norms(pb200_solver *s, int phase, const double *u_ana, double p, int relative, double out[4]);
/* Host-only helper (no device needed): coefficients of the polynomial preconditioner z = q(M^) r of the folded CG for the
 * Chebyshev interval [lo, hi]: step 1  z2 = out[0] r + out[2] M^ r ;  step 2  z3 = out[3] r + out[4] z2 + out[5] M^ z2  (out[1] = 0).       */
int pb200_poly_coefs(double lo, double hi, double out[6]);
int pb200_solver_destroy(pb200_solver *s);

#ifdef __cplusplus
}
#endif
#endif
