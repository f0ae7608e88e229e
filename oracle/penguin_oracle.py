"""CPU oracle for the Penguin.jl unsteady cut-cell diffusion hot path (operators + solver stages).

TEST INFRASTRUCTURE ONLY.  Nothing under ``penguin.jl_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use
it, and there only as the checker / the CPU arm that is timed beside the GPU.

It is a SciPy-sparse restatement of the reference's own assembly, block by block, so that the
matrix the reference would hand to UMFPACK is the matrix solved here:

* ``Mesh``                         -> /root/reference/src/mesh.jl:41-79
* ``delta_m`` / ``lift``           -> /root/reference/src/operators.jl:9, 92-113
* ``DiffusionOps``                 -> /root/reference/src/operators.jl:127-178
* ``grad`` / ``div``               -> /root/reference/src/operators.jl:20-34
* ``build_I_bc/I_D/source/g_g``    -> /root/reference/src/solver.jl:203-323
* ``BC_border_mono/diph``          -> /root/reference/src/solver.jl:379-580
* ``remove_zero_rows_cols``        -> /root/reference/src/solver.jl:59-78
* ``solve_system``                 -> /root/reference/src/solver.jl:158-188 (direct route, ``method = \\``)
* ``A_*/b_*`` builders, ctors and time loops -> /root/reference/src/solver/diffusion.jl:14-454

Parity status: the operator/solver stages are fully specified by in-tree Julia and are pinned
against the reference's own asserted values (tests/test_oracle_pins.py: ``max u1 = 1.15 +- 0.01``
from test/solver/diffusion_test.jl:54, ``max u_gamma = 1 +- 0.01`` from :80, the manufactured
solutions of test/convergence_test.jl:27,48,69, the no-body known answer of SURVEY Appendix A.1).
The geometry stage (capacities) lives in third-party code that is absent from /root/reference
(CartesianGeometry.jl 0.1.1 -> Vofinit.jl 0.1.0 -> libvofi 2.0.0); see oracle/geom_oracle.c whose
header says "parity unpinned" for the per-cell 1e-12 level.

Conventions: every per-cell array has the reference's padded length n = prod(n_i + 1),
x fastest (``idx = i + (n1+1) j + (n1+1)(n2+1) k``, 0-based here).
"""
from __future__ import annotations

import inspect
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------------
# Mesh  (src/mesh.jl:41-79)
# --------------------------------------------------------------------------------------------
class Mesh:
    """Uniform Cartesian mesh.  ``centers[d][j] = x0 + j h``, ``nodes[d][j] = x0 + (j + 0.5) h``."""

    def __init__(self, n, domain_size, x0=None):
        self.N = len(n)
        self.dims = tuple(int(v) for v in n)
        self.L = tuple(float(v) for v in domain_size)
        self.x0 = tuple(float(v) for v in (x0 if x0 is not None else (0.0,) * self.N))
        self.h = tuple(self.L[d] / self.dims[d] for d in range(self.N))
        # same floating-point expression as mesh.jl:49-50
        self.centers = tuple(np.array([self.x0[d] + j * (self.L[d] / self.dims[d]) for j in range(self.dims[d])])
                             for d in range(self.N))
        self.nodes = tuple(np.array([self.x0[d] + (j + 0.5) * (self.L[d] / self.dims[d]) for j in range(self.dims[d] + 1)])
                           for d in range(self.N))
        self.pdims = tuple(v + 1 for v in self.dims)          # padded dims
        self.n = int(np.prod(self.pdims))
        self._border = None

    @property
    def border_cells(self):
        """List of (0-based multi-index tuple) in the reference's order (mesh.jl:57-74): for each
        dim, for face in (first, last), product of the other ranges (first index fastest),
        de-duplicated keeping the first occurrence."""
        if self._border is None:
            seen, out = set(), []
            for d in range(self.N):
                for face in (0, self.dims[d] - 1):
                    ranges = [range(self.dims[i]) for i in range(self.N)]
                    ranges[d] = range(face, face + 1)
                    # Iterators.product: first iterator fastest
                    grids = np.meshgrid(*[np.array(r) for r in ranges], indexing="ij")
                    idx = np.stack([g.ravel(order="F") for g in grids], axis=1)
                    for row in idx:
                        t = tuple(int(v) for v in row)
                        if t not in seen:
                            seen.add(t)
                            out.append(t)
            self._border = out
        return self._border

    def lin(self, idx):
        """0-based linear index on the padded grid (solver.jl:362-372)."""
        li, stride = 0, 1
        for d in range(self.N):
            li += idx[d] * stride
            stride *= self.pdims[d]
        return li


# --------------------------------------------------------------------------------------------
# Capacity container (src/capacity.jl:25-36) -- arrays only, geometry comes from geom_oracle / imports
# --------------------------------------------------------------------------------------------
class Capacity:
    def __init__(self, mesh, V, Gamma, cell_types, A, B, W, C_omega, C_gamma=None):
        self.mesh = mesh
        self.N = mesh.N
        self.V = np.asarray(V, float)
        self.Gamma = np.asarray(Gamma, float)
        self.cell_types = np.asarray(cell_types, float)
        self.A = tuple(np.asarray(a, float) for a in A)
        self.B = tuple(np.asarray(b, float) for b in B)
        self.W = tuple(np.asarray(w, float) for w in W)
        self.C_omega = np.asarray(C_omega, float)           # (n, N)
        self.C_gamma = None if C_gamma is None else np.asarray(C_gamma, float)


def nobody_capacity(mesh):
    """Known answer of SURVEY Appendix A.1 (``body == -1``, examples/2D/Diffusion/Heat_Nobody.jl:13):
    every real cell full, W zero on the first/last face index of each direction."""
    N, pd = mesh.N, mesh.pdims
    shape = pd[::-1]                                       # C-order array with x fastest
    real = np.ones(shape, bool)
    for d in range(N):
        sl = [slice(None)] * N
        sl[N - 1 - d] = pd[d] - 1
        real[tuple(sl)] = False
    hN = float(np.prod(mesh.h))
    V = np.where(real, hN, 0.0).ravel()
    ct = np.where(real, 1.0, 0.0).ravel()
    A, B, W = [], [], []
    for d in range(N):
        face = hN / mesh.h[d]
        # A_d defined for face index 0..n_d in direction d, real cells in the other directions
        okA = np.ones(shape, bool)
        for e in range(N):
            if e != d:
                sl = [slice(None)] * N
                sl[N - 1 - e] = pd[e] - 1
                okA[tuple(sl)] = False
        A.append(np.where(okA, face, 0.0).ravel())
        B.append(np.where(real, face, 0.0).ravel())
        okW = real.copy()
        sl = [slice(None)] * N
        sl[N - 1 - d] = 0
        okW[tuple(sl)] = False
        W.append(np.where(okW, hN, 0.0).ravel())
    grids = np.meshgrid(*[mesh.nodes[d] + 0.5 * mesh.h[d] for d in range(N)], indexing="ij")
    C = np.stack([np.where(real, np.transpose(g, range(N)[::-1]), 0.0).ravel() for g in grids], axis=1)
    return Capacity(mesh, V, np.zeros(mesh.n), ct, A, B, W, C, np.zeros((mesh.n, N)))


# --------------------------------------------------------------------------------------------
# Elementary operators (src/operators.jl:9-13, 92-113)
# --------------------------------------------------------------------------------------------
def delta_m(n):
    """Backward difference with the reference's quirk: ``D[n, n] = 0`` (operators.jl:9)."""
    D = sp.diags([np.ones(n), -np.ones(n - 1)], [0, -1], format="csr")
    # Julia's `D[n, n] = 0.0` on a SparseMatrixCSC keeps the entry STORED with value zero (it matters only for NaN propagation,
    # solve_darcy_velocity: NaN * stored 0.0 = NaN) -- recalled SparseArrays behaviour, see SURVEY Appendix B
    D.data[D.indptr[n - 1]:D.indptr[n]][D.indices[D.indptr[n - 1]:D.indptr[n]] == n - 1] = 0.0
    return D


def lift(op1d_list):
    """``kron(op_N, ..., op_1)`` -- x fastest (operators.jl:104-110)."""
    res = op1d_list[-1]
    for op in op1d_list[-2::-1]:
        res = sp.kron(res, op, format="csr")
    return sp.csr_matrix(res)


class DiffusionOps:
    """G, H, W-dagger, V, size  (operators.jl:127-178)."""

    def __init__(self, cap: Capacity):
        mesh = cap.mesh
        N, pd = mesh.N, mesh.pdims
        self.cap = cap
        self.size = pd
        self.n = mesh.n
        Dm = []
        for d in range(N):
            ops = [delta_m(pd[i]) if i == d else sp.identity(pd[i], format="csr") for i in range(N)]
            Dm.append(lift(ops) if N > 1 else ops[0])
        self.Dm = Dm
        self.G = sp.vstack([Dm[d] @ sp.diags(cap.B[d]) for d in range(N)], format="csr")
        self.H = sp.vstack([sp.diags(cap.A[d]) @ Dm[d] - Dm[d] @ sp.diags(cap.B[d]) for d in range(N)], format="csr")
        w = np.concatenate(cap.W)
        wd = np.ones_like(w)
        nz = w != 0
        wd[nz] = 1.0 / w[nz]                               # 1.0 (not 0) where W == 0  (operators.jl:148-150)
        self.Wdag_diag = wd
        self.Wdag = sp.diags(wd, format="csr")
        self.V = sp.diags(cap.V, format="csr")


def delta_p(n):
    """Forward difference, last row zero (operators.jl:10)."""
    D = sp.lil_matrix(sp.diags([-np.ones(n), np.ones(n - 1)], [0, 1]))
    D[n - 1, n - 1] = 0.0
    return sp.csr_matrix(D)


def sigma_m(n):
    """Backward average; ``D[n, n] = 0`` (operators.jl:11)."""
    D = sp.lil_matrix(0.5 * sp.diags([np.ones(n), np.ones(n - 1)], [0, -1]))
    D[n - 1, n - 1] = 0.0
    return sp.csr_matrix(D)


def sigma_p(n):
    """Forward average, last row zero (operators.jl:12)."""
    D = sp.lil_matrix(0.5 * sp.diags([np.ones(n), np.ones(n - 1)], [0, 1]))
    D[n - 1, n - 1] = 0.0
    return sp.csr_matrix(D)


class ConvectionOps(DiffusionOps):
    """``ConvectionOps(capacity, u_omega, u_gamma)`` (operators.jl:194-209): C_d = D_p diag(S_m A_d u_omega_d) S_m,  K_d = diag(S_p H' u_gamma).
    ``uo``: N vectors of n bulk velocity components, ``ug``: one vector of N n interface velocity components."""

    def __init__(self, cap: Capacity, uo, ug):
        super().__init__(cap)
        mesh = cap.mesh
        N, pd = mesh.N, mesh.pdims

        def lifted(fn, d):
            ops = [fn(pd[i]) if i == d else sp.identity(pd[i], format="csr") for i in range(N)]
            return lift(ops) if N > 1 else ops[0]
        Dp = [lifted(delta_p, d) for d in range(N)]
        Sm = [lifted(sigma_m, d) for d in range(N)]
        Sp = [lifted(sigma_p, d) for d in range(N)]
        ug = np.asarray(ug, float)
        self.C = [Dp[d] @ sp.diags(Sm[d] @ (cap.A[d] * np.asarray(uo[d], float))) @ Sm[d] for d in range(N)]
        q = self.H.T @ ug
        self.K = [sp.diags(Sp[d] @ q) for d in range(N)]


def _conv(op):
    return sum(op.C[1:], op.C[0]), 0.5 * sum(op.K[1:], op.K[0])


def A_mono_stead_advdiff(op, cap, D, bc):
    """advectiondiffusion.jl:30-47"""
    ia, ib = build_I_bc(op.n, bc)
    Id = build_I_D(op, D, cap)
    GG, GH, HG, HH = _blocks(op, Id)
    Cb, Ki = _conv(op)
    Ib, Ia, Ig = sp.diags(ib), sp.diags(ia), sp.diags(cap.Gamma)
    return sp.bmat([[Cb + Ki + GG, Ki + GH], [Ib @ HG, Ib @ HH + Ia @ Ig]], format="csr")


def A_mono_unstead_advdiff(op, cap, D, bc, dt, scheme):
    """advectiondiffusion.jl:178-213"""
    ia, ib = build_I_bc(op.n, bc)
    Id = build_I_D(op, D, cap)
    GG, GH, HG, HH = _blocks(op, Id)
    Cb, Ki = _conv(op)
    Ib, Ia, Ig = sp.diags(ib), sp.diags(ia), sp.diags(cap.Gamma)
    tie = Ib @ HH + Ia @ Ig
    if scheme == "CN":
        return sp.bmat([[op.V + dt / 2 * (Cb + Ki + GG), dt / 2 * (Ki + GH)], [dt / 2 * (Ib @ HG), dt / 2 * tie]], format="csr")
    return sp.bmat([[op.V + dt * (Cb + Ki + GG), dt * (Ki + GH)], [Ib @ HG, tie]], format="csr")


def b_mono_unstead_advdiff(op, f, cap, D, bc, Ti, dt, t, scheme):
    """advectiondiffusion.jl:215-252"""
    n = op.n
    fn, fn1 = build_source(op, f, cap, t), build_source(op, f, cap, t + dt)
    gn, gn1 = build_g_g(op, bc, cap, t), build_g_g(op, bc, cap, t + dt)
    ia, ib = build_I_bc(n, bc)
    Id = build_I_D(op, D, cap)
    To, Tg = Ti[:n], Ti[n:]
    V, Gam = cap.V, cap.Gamma
    if scheme == "CN":
        GG, GH, HG, HH = _blocks(op, Id)
        Cb, Ki = _conv(op)
        b1 = V * To - dt / 2 * ((Cb + Ki + GG) @ To) - dt / 2 * ((Ki + GH) @ Tg) + dt / 2 * V * (fn + fn1)
        b2 = dt / 2 * Gam * (gn + gn1) - dt / 2 * ib * (HG @ To) - dt / 2 * (ib * (HH @ Tg) + ia * Gam * Tg)
    else:
        b1 = V * To + dt * V * fn1
        b2 = Gam * gn1
    return np.concatenate([b1, b2])


def AdvectionDiffusionSteadyMono(phase, bc_b, bc_i):
    """advectiondiffusion.jl:12-28"""
    s = Solver()
    s.A = A_mono_stead_advdiff(phase.operator, phase.capacity, phase.D, bc_i)
    s.b = b_mono_stead_diff(phase.operator, phase.source, phase.capacity, bc_i)      # (b_mono_stead_advdiff :49-63 is the same vector)
    s.A, s.b = BC_border_mono(s.A, s.b, bc_b, phase.capacity.mesh)
    return s


def solve_AdvectionDiffusionSteadyMono(s):
    s.x = solve_system(s.A, s.b)
    return s


def AdvectionDiffusionUnsteadyMono(phase, bc_b, bc_i, dt, Ti, scheme):
    """advectiondiffusion.jl:163-176: NO border rows in the constructor's system (unlike the diffusion constructor)."""
    s = Solver()
    sch = "CN" if scheme == "CN" else "BE"
    s.A = A_mono_unstead_advdiff(phase.operator, phase.capacity, phase.D, bc_i, dt, sch)
    s.b = b_mono_unstead_advdiff(phase.operator, phase.source, phase.capacity, phase.D, bc_i, np.asarray(Ti, float), dt, 0.0, sch)
    return s


def solve_AdvectionDiffusionUnsteadyMono(s, phase, dt, Tend, bc_b, bc, scheme, max_steps=None):
    """advectiondiffusion.jl:254-283.  The reference's loop calls b_mono_unstead_advdiff WITHOUT the diffusion coefficient (:272, a MethodError as
    written); this restates the evident intent (the 9-argument call of the constructor, :175)."""
    t = 0.0
    s.x = solve_system(s.A, s.b)
    s.states.append(s.x)
    k = 0
    while t < Tend and (max_steps is None or k < max_steps):
        t += dt
        s.A = A_mono_unstead_advdiff(phase.operator, phase.capacity, phase.D, bc, dt, scheme)
        s.b = b_mono_unstead_advdiff(phase.operator, phase.source, phase.capacity, phase.D, bc, s.x, dt, t, scheme)
        s.A, s.b = BC_border_mono(s.A, s.b, bc_b, phase.capacity.mesh, t=t)
        s.x = solve_system(s.A, s.b)
        s.states.append(s.x)
        k += 1
    return s


def grad(op: DiffusionOps, p):
    """operators.jl:20-23"""
    n = op.n
    return op.Wdag @ (op.G @ p[:n] + op.H @ p[n:])


def div(op: DiffusionOps, q_omega, q_gamma):
    """operators.jl:30-34"""
    GT, HT = op.G.T, op.H.T
    return -(GT + HT) @ q_omega + HT @ q_gamma


# --------------------------------------------------------------------------------------------
# Boundary-condition value types (src/boundary.jl)
# --------------------------------------------------------------------------------------------
class Dirichlet:
    def __init__(self, value): self.value = value


class Neumann:
    def __init__(self, value): self.value = value


class Robin:
    def __init__(self, alpha, beta, value): self.alpha, self.beta, self.value = alpha, beta, value


class Periodic:
    pass


class ScalarJump:
    def __init__(self, a1, a2, value): self.a1, self.a2, self.value = a1, a2, value


class FluxJump:
    def __init__(self, b1, b2, value): self.b1, self.b2, self.value = b1, b2, value


class InterfaceConditions:
    def __init__(self, scalar, flux): self.scalar, self.flux = scalar, flux


class BorderConditions:
    def __init__(self, borders=None): self.borders = dict(borders or {})


class Phase:
    def __init__(self, capacity, operator, source, D):
        self.capacity, self.operator, self.source, self.D = capacity, operator, source, D


# --------------------------------------------------------------------------------------------
# Coefficient / RHS helpers  (src/solver.jl:203-323)
# --------------------------------------------------------------------------------------------
def _nargs(f):
    try:
        return len([p for p in inspect.signature(f).parameters.values()
                    if p.default is inspect.Parameter.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
    except (TypeError, ValueError):
        return None


def _coords3(C):
    """get_all_coordinates (solver.jl:230-248): pad to (x, y, z) with zeros."""
    n, N = C.shape
    cols = [C[:, d] for d in range(N)] + [np.zeros(n)] * (3 - N)
    return cols[:3] if N <= 3 else [C[:, d] for d in range(N)]


def _eval(f, cols, t=None, n=None):
    """Evaluate a (vectorised) closure the way the reference calls it; constants are broadcast."""
    if not callable(f):
        return np.full(n, float(f))
    if t is None:
        out = f(*cols)
    else:
        try:
            out = f(*cols, t)                              # time-dependent first (solver.jl:315-319)
        except TypeError:
            out = f(*cols)
    return np.broadcast_to(np.asarray(out, float), (n,)).copy()


def build_I_bc(n, bc):
    """(I_alpha, I_beta) as diagonal vectors (solver.jl:203-223)."""
    ia, ib = np.zeros(n), np.zeros(n)
    if isinstance(bc, Dirichlet):
        ia[:] = 1.0
    elif isinstance(bc, Neumann):
        ib[:] = 1.0
    elif isinstance(bc, Robin):
        ia[:] = bc.alpha
        ib[:] = bc.beta
    return ia, ib


def build_I_D(op, D, cap):
    return _eval(D, _coords3(cap.C_omega), None, op.n)


def build_source(op, f, cap, t=None):
    return _eval(f, _coords3(cap.C_omega), t, op.n)


def build_g_g(op, bc, cap, t=None):
    if callable(bc.value):
        return _eval(bc.value, _coords3(cap.C_gamma), t, op.n)
    return np.full(op.n, float(bc.value))


# --------------------------------------------------------------------------------------------
# Border conditions (src/solver.jl:379-580)
# --------------------------------------------------------------------------------------------
def classify_boundary_cell_fast(ci, mesh):
    """solver.jl:379-409 -- note :left/:right are the extremes of grid dim 2 (index 1 here)."""
    N = mesh.N
    if N >= 2:
        if ci[1] == 0:
            return "left"
        if ci[1] == mesh.dims[1] - 1:
            return "right"
    if ci[0] == 0:
        return "bottom"
    if ci[0] == mesh.dims[0] - 1:
        return "top"
    if N >= 3:
        if ci[2] == 0:
            return "backward"
        if ci[2] == mesh.dims[2] - 1:
            return "forward"
    raise RuntimeError("cell not on a border")


_OPP = {"left": "right", "right": "left", "bottom": "top", "top": "bottom", "backward": "forward", "forward": "backward"}


def _eval_bc_value(value, pos, t):
    """eval_bc_value (solver.jl:441-448): pos has N entries (mesh.centers coordinates)."""
    if not callable(value):
        return float(value)
    if t is None:
        return float(value(*pos))
    try:
        return float(value(*pos, t))
    except TypeError:
        return float(value(*pos))


def _periodic_partner(ci, key, mesh):
    """find_corresponding_cell_optimized (solver.jl:506-530): maps to the PADDED extreme."""
    pd = mesh.pdims
    c = list(ci)
    if key == "left":
        c[1] = pd[1] - 1
    elif key == "right":
        c[1] = 0
    elif key == "bottom":
        c[0] = pd[0] - 1
    elif key == "top":
        c[0] = 0
    elif key == "backward":
        c[2] = pd[2] - 1
    elif key == "forward":
        c[2] = 0
    return mesh.lin(c)


def _apply_bc(A, b, li, pos, cond, key, bc_b, mesh, t, offset=0):
    """apply_boundary_condition_fast! (solver.jl:450-499). ``A`` is LIL."""
    row = li + offset
    if isinstance(cond, Dirichlet):
        A.rows[row] = [row]
        A.data[row] = [1.0]
        b[row] = _eval_bc_value(cond.value, pos, t)
    elif isinstance(cond, Periodic):
        if _OPP[key] in bc_b.borders:
            ci = np.unravel_index(li, mesh.pdims[::-1])[::-1]
            other = _periodic_partner(ci, key, mesh) + offset
            A.rows[row] = []
            A.data[row] = []
            A[row, row] = A[row, row] + 1.0
            A[row, other] = A[row, other] - 1.0
            b[row] = 0.0
    elif isinstance(cond, Neumann):
        if mesh.N == 1:
            dx = float(np.min(np.diff(mesh.nodes[0])))
            nn = mesh.pdims[0]
            adj = min(li + 1, nn - 1) if key == "bottom" else max(li - 1, 0)
            g = _eval_bc_value(cond.value, pos, t)
            A.rows[row] = []
            A.data[row] = []
            A[row, row] = 1.0 / dx
            A[row, adj + offset] = A[row, adj + offset] - 1.0 / dx
            b[row] = g
        # >= 2-D: warning no-op (solver.jl:494-496)
    # Robin / others: silently nothing


def BC_border_mono(A, b, bc_b, mesh, t=None):
    A = A.tolil()
    for ci in mesh.border_cells:
        key = classify_boundary_cell_fast(ci, mesh)
        cond = bc_b.borders.get(key)
        if cond is None:
            continue
        pos = tuple(mesh.centers[d][ci[d]] for d in range(mesh.N))
        _apply_bc(A, b, mesh.lin(ci), pos, cond, key, bc_b, mesh, t)
    return A.tocsr(), b


def BC_border_diph(A, b, bc_b, cap1, cap2, t=None):
    mesh = cap1.mesh
    n = A.shape[0] // 4
    A = A.tolil()
    for ci in mesh.border_cells:
        key = classify_boundary_cell_fast(ci, mesh)
        cond = bc_b.borders.get(key)
        if cond is None:
            continue
        li = mesh.lin(ci)
        pos = tuple(mesh.centers[d][ci[d]] for d in range(mesh.N))
        for off, cap in ((0, cap1), (2 * n, cap2)):
            if cap.cell_types[li] == 0:
                continue                                   # solver.jl:574-575
            _apply_bc(A, b, li, pos, cond, key, bc_b, mesh, t, offset=off)
    return A.tocsr(), b


# --------------------------------------------------------------------------------------------
# Linear solve (src/solver.jl:59-78, 158-188)
# --------------------------------------------------------------------------------------------
def remove_zero_rows_cols(A):
    absA = abs(A)
    rs = np.asarray(absA.sum(axis=1)).ravel()
    cs = np.asarray(absA.sum(axis=0)).ravel()
    return np.nonzero((rs != 0.0) & (cs != 0.0))[0]


def solve_system(A, b):
    """Direct route of solve_system! (``method = \\``): trim, LU, scatter into zeros(n)."""
    A = sp.csr_matrix(A)
    idx = remove_zero_rows_cols(A)
    Ar = A[idx][:, idx].tocsc()
    lu = spla.splu(Ar)
    br = b[idx]
    xr = lu.solve(br)
    # Julia's `\` on a sparse square matrix is UMFPACK, which refines the solution (UMFPACK_IRSTEP = 2 by default); SuperLU does not.
    # Without it two orderings of the same LU differ by 2e-9 on a 3-D diphasic system with near-empty cut cells -- above the 1e-9 bar.
    for _ in range(2):
        xr = xr + lu.solve(br - Ar @ xr)
    x = np.zeros(A.shape[0])
    x[idx] = xr
    return x


class Solver:
    def __init__(self):
        self.A = None
        self.b = None
        self.x = None
        self.states = []
        self.ch = []


# --------------------------------------------------------------------------------------------
# Block systems (src/solver/diffusion.jl)
# --------------------------------------------------------------------------------------------
def _blocks(op, Id):
    G, H, W = op.G, op.H, op.Wdag
    GT, HT = G.T.tocsr(), H.T.tocsr()
    D = sp.diags(Id)
    return D @ GT @ W @ G, D @ GT @ W @ H, HT @ W @ G, HT @ W @ H


def A_mono_stead_diff(op, cap, D, bc):
    """diffusion.jl:30-43"""
    ia, ib = build_I_bc(op.n, bc)
    Id = build_I_D(op, D, cap)
    GG, GH, HG, HH = _blocks(op, Id)
    Ib, Ia, Ig = sp.diags(ib), sp.diags(ia), sp.diags(cap.Gamma)
    return sp.bmat([[GG, GH], [Ib @ HG, Ib @ HH + Ia @ Ig]], format="csr")


def b_mono_stead_diff(op, f, cap, bc):
    """diffusion.jl:45-58"""
    fo = build_source(op, f, cap)
    gg = build_g_g(op, bc, cap)
    return np.concatenate([cap.V * fo, cap.Gamma * gg])


def A_mono_unstead_diff(op, cap, D, bc, dt, scheme):
    """diffusion.jl:212-241"""
    ia, ib = build_I_bc(op.n, bc)
    Id = build_I_D(op, D, cap)
    GG, GH, HG, HH = _blocks(op, Id)
    Ib, Ia, Ig = sp.diags(ib), sp.diags(ia), sp.diags(cap.Gamma)
    if scheme == "CN":
        b1 = op.V + dt / 2 * GG
        b2 = dt / 2 * GH
        b3 = dt / 2 * (Ib @ HG)
        b4 = dt / 2 * (Ib @ HH) + dt / 2 * (Ia @ Ig)
    else:
        b1 = op.V + dt * GG
        b2 = dt * GH
        b3 = Ib @ HG
        b4 = Ib @ HH + Ia @ Ig
    return sp.bmat([[b1, b2], [b3, b4]], format="csr")


def b_mono_unstead_diff(op, f, D, cap, bc, Ti, dt, t, scheme):
    """diffusion.jl:243-265"""
    n = op.n
    fn, fn1 = build_source(op, f, cap, t), build_source(op, f, cap, t + dt)
    gn, gn1 = build_g_g(op, bc, cap, t), build_g_g(op, bc, cap, t + dt)
    ia, ib = build_I_bc(n, bc)
    Id = build_I_D(op, D, cap)
    To, Tg = Ti[:n], Ti[n:]
    V, Gam = cap.V, cap.Gamma
    if scheme == "CN":
        GG, GH, HG, HH = _blocks(op, Id)
        b1 = V * To - dt / 2 * (GG @ To) - dt / 2 * (GH @ Tg) + dt / 2 * V * (fn + fn1)
        b2 = dt / 2 * Gam * (gn + gn1) - dt / 2 * ib * (HG @ To) - dt / 2 * ib * (HH @ Tg) - dt / 2 * ia * Gam * Tg
    else:
        b1 = V * To + dt * V * fn1
        b2 = Gam * gn1
    return np.concatenate([b1, b2])


def A_diph_stead_diff(op1, op2, cap1, cap2, D1, D2, ic):
    """diffusion.jl:104-144"""
    n = op1.n
    a1, a2 = float(ic.scalar.a1), float(ic.scalar.a2)
    be1, be2 = float(ic.flux.b1), float(ic.flux.b2)
    GG1, GH1, HG1, HH1 = _blocks(op1, build_I_D(op1, D1, cap1))
    GG2, GH2, HG2, HH2 = _blocks(op2, build_I_D(op2, D2, cap2))
    I = sp.identity(n, format="csr")
    Z = sp.csr_matrix((n, n))
    return sp.bmat([[GG1, GH1, Z, Z],
                    [Z, a1 * I, Z, -a2 * I],
                    [Z, Z, GG2, GH2],
                    [be1 * HG1, be1 * HH1, be2 * HG2, be2 * HH2]], format="csr")


def b_diph_stead_diff(op1, op2, f1, f2, cap1, cap2, ic):
    """diffusion.jl:146-161"""
    g = build_g_g(op1, ic.scalar, cap1)
    h = build_g_g(op2, ic.flux, cap2)
    return np.concatenate([cap1.V * build_source(op1, f1, cap1), g, cap2.V * build_source(op2, f2, cap2), cap2.Gamma * h])


def A_diph_unstead_diff(op1, op2, cap1, cap2, D1, D2, ic, dt, scheme):
    """diffusion.jl:334-389"""
    n = op1.n
    a1, a2 = float(ic.scalar.a1), float(ic.scalar.a2)
    be1, be2 = float(ic.flux.b1), float(ic.flux.b2)
    GG1, GH1, HG1, HH1 = _blocks(op1, build_I_D(op1, D1, cap1))
    GG2, GH2, HG2, HH2 = _blocks(op2, build_I_D(op2, D2, cap2))
    c = dt / 2 if scheme == "CN" else dt
    I = sp.identity(n, format="csr")
    Z = sp.csr_matrix((n, n))
    return sp.bmat([[op1.V + c * GG1, c * GH1, Z, Z],
                    [Z, a1 * I, Z, -a2 * I],
                    [Z, Z, op2.V + c * GG2, c * GH2],
                    [be1 * HG1, be1 * HH1, be2 * HG2, be2 * HH2]], format="csr")


def b_diph_unstead_diff(op1, op2, f1, f2, cap1, cap2, D1, D2, ic, Ti, dt, t, scheme):
    """diffusion.jl:391-420"""
    n = op1.n
    g = build_g_g(op1, ic.scalar, cap1)
    h = build_g_g(op2, ic.flux, cap2)
    f1n, f2n = build_source(op1, f1, cap1, t), build_source(op2, f2, cap2, t)
    f1p, f2p = build_source(op1, f1, cap1, t + dt), build_source(op2, f2, cap2, t + dt)
    To1, Tg1, To2, Tg2 = Ti[:n], Ti[n:2 * n], Ti[2 * n:3 * n], Ti[3 * n:]
    if scheme == "CN":
        GG1, GH1, _, _ = _blocks(op1, build_I_D(op1, D1, cap1))
        GG2, GH2, _, _ = _blocks(op2, build_I_D(op2, D2, cap2))
        b1 = cap1.V * To1 - dt / 2 * (GG1 @ To1) - dt / 2 * (GH1 @ Tg1) + dt / 2 * cap1.V * (f1n + f1p)
        b3 = cap2.V * To2 - dt / 2 * (GG2 @ To2) - dt / 2 * (GH2 @ Tg2) + dt / 2 * cap2.V * (f2n + f2p)
    else:
        b1 = cap1.V * To1 + dt * cap1.V * f1p
        b3 = cap2.V * To2 + dt * cap2.V * f2p
    return np.concatenate([b1, g, b3, cap2.Gamma * h])


# --------------------------------------------------------------------------------------------
# Constructors + time loops (src/solver/diffusion.jl:14-28, 60-72, 88-102, 192-210, 268-301, 319-332, 422-454)
# --------------------------------------------------------------------------------------------
def DiffusionSteadyMono(phase, bc_b, bc_i):
    s = Solver()
    s.A = A_mono_stead_diff(phase.operator, phase.capacity, phase.D, bc_i)
    s.b = b_mono_stead_diff(phase.operator, phase.source, phase.capacity, bc_i)
    s.A, s.b = BC_border_mono(s.A, s.b, bc_b, phase.capacity.mesh)
    return s


def solve_DiffusionSteadyMono(s):
    s.x = solve_system(s.A, s.b)
    return s


def DiffusionSteadyDiph(ph1, ph2, bc_b, ic):
    s = Solver()
    s.A = A_diph_stead_diff(ph1.operator, ph2.operator, ph1.capacity, ph2.capacity, ph1.D, ph2.D, ic)
    s.b = b_diph_stead_diff(ph1.operator, ph2.operator, ph1.source, ph2.source, ph1.capacity, ph2.capacity, ic)
    s.A, s.b = BC_border_diph(s.A, s.b, bc_b, ph1.capacity, ph2.capacity)
    return s


def solve_DiffusionSteadyDiph(s):
    s.x = solve_system(s.A, s.b)
    return s


def DiffusionUnsteadyMono(phase, bc_b, bc_i, dt, Ti, scheme):
    s = Solver()
    sch = "CN" if scheme == "CN" else "BE"
    s.A = A_mono_unstead_diff(phase.operator, phase.capacity, phase.D, bc_i, dt, sch)
    s.b = b_mono_unstead_diff(phase.operator, phase.source, phase.D, phase.capacity, bc_i, np.asarray(Ti, float), dt, 0.0, sch)
    s.A, s.b = BC_border_mono(s.A, s.b, bc_b, phase.capacity.mesh, t=0.0)
    return s


def solve_DiffusionUnsteadyMono(s, phase, dt, Tend, bc_b, bc, scheme, max_steps=None):
    t = 0.0
    s.x = solve_system(s.A, s.b)
    s.states.append(s.x)
    Ti = s.x
    s.A = A_mono_unstead_diff(phase.operator, phase.capacity, phase.D, bc, dt, scheme)
    k = 0
    while t < Tend and (max_steps is None or k < max_steps):
        t += dt
        s.b = b_mono_unstead_diff(phase.operator, phase.source, phase.D, phase.capacity, bc, Ti, dt, t, scheme)
        s.A, s.b = BC_border_mono(s.A, s.b, bc_b, phase.capacity.mesh, t=t)
        s.x = solve_system(s.A, s.b)
        s.states.append(s.x)
        Ti = s.x
        k += 1
    return s


def DiffusionUnsteadyDiph(ph1, ph2, bc_b, ic, dt, Ti, scheme):
    s = Solver()
    s.A = A_diph_unstead_diff(ph1.operator, ph2.operator, ph1.capacity, ph2.capacity, ph1.D, ph2.D, ic, dt, scheme)
    s.b = b_diph_unstead_diff(ph1.operator, ph2.operator, ph1.source, ph2.source, ph1.capacity, ph2.capacity,
                              ph1.D, ph2.D, ic, np.asarray(Ti, float), dt, 0.0, scheme)
    s.A, s.b = BC_border_diph(s.A, s.b, bc_b, ph1.capacity, ph2.capacity)       # no t (diffusion.jl:330)
    return s


def solve_DiffusionUnsteadyDiph(s, ph1, ph2, dt, Tend, bc_b, ic, scheme, max_steps=None):
    t = 0.0
    s.x = solve_system(s.A, s.b)
    s.states.append(s.x)
    Ti = s.x
    s.A = A_diph_unstead_diff(ph1.operator, ph2.operator, ph1.capacity, ph2.capacity, ph1.D, ph2.D, ic, dt, scheme)
    k = 0
    while t < Tend and (max_steps is None or k < max_steps):
        t += dt
        s.b = b_diph_unstead_diff(ph1.operator, ph2.operator, ph1.source, ph2.source, ph1.capacity, ph2.capacity,
                                  ph1.D, ph2.D, ic, Ti, dt, t, scheme)
        s.A, s.b = BC_border_diph(s.A, s.b, bc_b, ph1.capacity, ph2.capacity)   # no t (diffusion.jl:446)
        s.x = solve_system(s.A, s.b)
        s.states.append(s.x)
        Ti = s.x
        k += 1
    return s


def n_solves(dt, Tend):
    """Number of solves the reference loop performs: 1 + #{iterations of ``while t < Tend; t += dt``}."""
    t, k = 0.0, 0
    while t < Tend:
        t += dt
        k += 1
    return 1 + k


# --------------------------------------------------------------------------------------------
# check_convergence (src/convergence.jl:4-93) -- used only to pin the oracle on the reference's asserts
# --------------------------------------------------------------------------------------------
def check_convergence(u_analytical, x, cap, p=2, relative=False):
    C = cap.C_omega
    u_ana = np.asarray(u_analytical(*[C[:, d] for d in range(cap.N)]), float) * np.ones(len(C))
    u_num = x[:len(x) // 2] if len(x) == 2 * len(C) else x
    err = u_ana - u_num
    ct = cap.cell_types

    def lp(mask):
        if p == np.inf:
            if not relative:
                return float(np.max(np.abs(err[mask]), initial=0.0))
            # errors[idx] / u_ana[idx] on two Julia vectors is err * pinv(u_ana), a matrix (src/convergence.jl:19)
            with np.errstate(all="ignore"):
                return float(np.max(np.abs(err[mask]), initial=0.0) * np.max(np.abs(u_ana[mask]), initial=0.0) / np.sum(u_ana[mask] ** 2))
        with np.errstate(all="ignore"):
            e = np.abs(err[mask] / u_ana[mask]) if relative else np.abs(err[mask])
            return float((np.sum(e ** p * cap.V[mask]) / np.sum(cap.V)) ** (1.0 / p))
    return lp((ct == 1) | (ct == -1)), lp(ct == 1), lp(ct == -1), lp(ct == 0)


# --------------------------------------------------------------------------------------------
# Darcy (src/solver/darcy.jl:1-40): steady diffusion system + u = -grad p with NaN-masked pressures
# --------------------------------------------------------------------------------------------
def solve_darcy_velocity(x, op, cap):
    """u = -grad(op, p) with p_omega = NaN on empty cells and p_gamma = NaN on empty and full cells (darcy.jl:26-40).
    NaN propagation follows Julia's STRUCTURAL sparsity (recalled SparseArrays behaviour, SURVEY Appendix B: `spdiagm`, `D[n, n] = 0.0`,
    `kron`, `*` and `-` all keep numerically zero entries stored, and NaN * stored 0.0 = NaN), whereas SciPy prunes zeros while it builds
    G and H.  The values therefore come from the pruned operators applied to the NaN-free pressures, and an entry is NaN exactly when
    a structurally present coefficient of its row -- (i, i) and (i, i - 1) of the lifted backward difference, both in G and in H --
    meets a NaN."""
    n = len(x) // 2
    po, pg = np.array(x[:n], float), np.array(x[n:], float)
    ct = cap.cell_types
    bad_o = ct == 0
    bad_g = (ct == 0) | (ct == 1)
    po[bad_o] = 0.0
    pg[bad_g] = 0.0
    u = -grad(op, np.concatenate([po, pg]))
    bad = (bad_o | bad_g).astype(float)
    dims = cap.mesh.dims
    N = len(dims)
    for d in range(N):
        ops = []
        for i in range(N):
            m = dims[i] + 1
            ops.append(sp.diags([np.ones(m), np.ones(m - 1)], [0, -1], format="csr") if i == d else sp.identity(m, format="csr"))
        P = lift(ops)
        u[d * n:(d + 1) * n][(P @ bad) > 0] = np.nan
    return u
