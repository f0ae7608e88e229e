"""Host-side mirror of Penguin.jl's exported API for the unsteady cut-cell diffusion path, over libpenguin_b200.so.

The reference is Julia and Julia is not in this image, so this Python layer plays the role of the Julia shim shown in
INTEGRATION.md: same names, argument order and semantics as /root/reference/src (cited per symbol), every number
computed by the CUDA library through its C ABI.  Julia's ``f!`` is spelled ``f_`` here.

Differences forced by the device (hard part 2): ``body`` must be a GPU-evaluable level-set
descriptor (``Circle``/``Sphere``/``Balls``/``HalfSpace``, negated with unary minus like the reference's
``-(...)`` bodies) -- arbitrary closures go through ``Capacity.from_arrays``; ``f``, ``D`` and boundary values
remain arbitrary (vectorised) Python callables, evaluated on the host at ``C_ω`` / ``C_γ`` exactly where the
reference evaluates them (src/solver.jl:230-323).
"""
from __future__ import annotations

import atexit
import ctypes as C
import os

import numpy as np

from . import _lib as L

__all__ = [
    "init", "init_distributed", "init_multi", "finalize", "context", "Mesh", "Circle", "Sphere", "Interval", "Balls", "HalfSpace", "Capacity",
    "DiffusionOps", "grad", "div", "Phase", "Dirichlet", "Neumann", "Robin", "Periodic", "ScalarJump", "FluxJump",
    "BorderConditions", "InterfaceConditions", "Solver", "DiffusionSteadyMono", "solve_DiffusionSteadyMono_",
    "DiffusionSteadyDiph", "solve_DiffusionSteadyDiph_", "DiffusionUnsteadyMono", "solve_DiffusionUnsteadyMono_",
    "DiffusionUnsteadyDiph", "solve_DiffusionUnsteadyDiph_", "Steady", "Unsteady", "Monophasic", "Diphasic", "Diffusion",
    "DarcyFlow", "solve_DarcyFlow_", "DarcyFlowUnsteady", "solve_DarcyFlowUnsteady_", "solve_darcy_velocity", "check_convergence",
    "ConvectionOps", "DiffusionAdvection", "AdvectionDiffusionSteadyMono", "solve_AdvectionDiffusionSteadyMono_",
    "AdvectionDiffusionUnsteadyMono", "solve_AdvectionDiffusionUnsteadyMono_",
]

Steady, Unsteady = "Steady", "Unsteady"
Monophasic, Diphasic = "Monophasic", "Diphasic"
Diffusion = "Diffusion"
DiffusionAdvection = "DiffusionAdvection"


# ----------------------------------------------------------------------------------------------------------------
# context
# ----------------------------------------------------------------------------------------------------------------
class _Context:
    def __init__(self, handle, rank=0, nranks=1):
        self.h, self.rank, self.nranks = handle, rank, nranks

    def sync(self):
        L.check(L.lib().pb200_sync(self.h), self.h)

    @property
    def launches(self):
        return int(L.lib().pb200_launch_count(self.h))

    @property
    def stream(self):
        return int(L.lib().pb200_stream(self.h))


_ctx = None


def init(device=None):
    """pb200_init: one context per process (single GPU)."""
    global _ctx
    if _ctx is None:
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "PB200_DEVICE" not in os.environ else int(os.environ["PB200_DEVICE"])
        h = C.c_void_p()
        L.check(L.lib().pb200_init(C.byref(h), int(device)))
        _ctx = _Context(h)
    return _ctx


def init_distributed(rank, nranks, device, bcast):
    """One process per GPU.  ``bcast(bytes_or_None) -> bytes`` broadcasts rank 0's 128-byte NCCL id (the host uses
    torch.distributed / any launcher for that; the library itself only needs the id)."""
    global _ctx
    if _ctx is not None:
        raise RuntimeError("context already initialised")
    ident = None
    if rank == 0:
        buf = C.create_string_buffer(128)
        L.check(L.lib().pb200_nccl_unique_id(buf))
        ident = buf.raw
    ident = bcast(ident)
    h = C.c_void_p()
    L.check(L.lib().pb200_init_dist(C.byref(h), int(device), int(rank), int(nranks), ident))
    _ctx = _Context(h, rank, nranks)
    return _ctx


def init_multi(devices):
    """pb200_init_multi: ONE process driving several GPUs.  Every per-cell array keeps the GLOBAL padded length; the library cuts the slabs."""
    global _ctx
    if _ctx is not None:
        raise RuntimeError("context already initialised")
    dev = (C.c_int * len(devices))(*[int(d) for d in devices])
    h = C.c_void_p()
    L.check(L.lib().pb200_init_multi(C.byref(h), C.cast(dev, L.ip), len(devices)))
    _ctx = _Context(h)
    return _ctx


def context():
    return _ctx if _ctx is not None else init()


@atexit.register
def _finalize_at_exit():
    # handles that are still alive when the interpreter shuts down must not call into a CUDA runtime that is being torn down:
    # finalising here (before module teardown) makes their __del__ a no-op
    try:
        finalize()
    except Exception:
        pass


def finalize():
    global _ctx
    if _ctx is not None:
        L.lib().pb200_finalize(_ctx.h)
        _ctx = None


def _dp(a):
    return None if a is None else a.ctypes.data_as(L.dp)


# ----------------------------------------------------------------------------------------------------------------
# Mesh  (src/mesh.jl:41-79)
# ----------------------------------------------------------------------------------------------------------------
class _MeshTag:
    def __init__(self, mesh):
        self._mesh = mesh
        self._cells = None

    @property
    def border_cells(self):
        """[(CartesianIndex (1-based tuple), centers-coords)] in the reference's order (mesh.jl:57-74)."""
        if self._cells is None:
            m = self._mesh
            seen, out = set(), []
            for d in range(m.N):
                for face in (0, m.dims[d] - 1):
                    ranges = [np.arange(m.dims[i]) for i in range(m.N)]
                    ranges[d] = np.array([face])
                    grids = np.meshgrid(*ranges, indexing="ij")
                    for row in np.stack([g.ravel(order="F") for g in grids], axis=1):
                        t = tuple(int(v) for v in row)
                        if t not in seen:
                            seen.add(t)
                            out.append((tuple(v + 1 for v in t), tuple(float(m.centers[i][t[i]]) for i in range(m.N))))
            self._cells = out
        return self._cells


class Mesh:
    """``Mesh(n, domain_size, x0)``: ``centers[d][j] = x0 + j h``, ``nodes[d][j] = x0 + (j + 1/2) h``."""

    def __init__(self, n, domain_size, x0=None):
        self.N = len(n)
        self.dims = tuple(int(v) for v in n)
        self.domain_size = tuple(float(v) for v in domain_size)
        self.x0 = tuple(float(v) for v in (x0 if x0 is not None else (0.0,) * self.N))
        self.centers = tuple(np.array([self.x0[d] + j * (self.domain_size[d] / self.dims[d]) for j in range(self.dims[d])])
                             for d in range(self.N))
        self.nodes = tuple(np.array([self.x0[d] + (j + 0.5) * (self.domain_size[d] / self.dims[d]) for j in range(self.dims[d] + 1)])
                           for d in range(self.N))
        self.tag = _MeshTag(self)

    def nC(self):
        return int(np.prod(self.dims))


# ----------------------------------------------------------------------------------------------------------------
# bodies: GPU-evaluable level sets that are also plain callables (so host code can still evaluate body(x, y))
# ----------------------------------------------------------------------------------------------------------------
class Balls:
    """phi(x) = min_k |x - c_k| - r_k for DISJOINT balls; ``-body`` is the sign flip."""
    kind = 0

    def __init__(self, centers, radii, fluid_inside=True):
        self.centers = np.atleast_2d(np.asarray(centers, float))
        self.radii = np.atleast_1d(np.asarray(radii, float))
        self.fluid_inside = bool(fluid_inside)
        assert self.centers.shape[0] == self.radii.shape[0]

    def __neg__(self):
        return Balls(self.centers, self.radii, not self.fluid_inside)

    def __call__(self, *x):
        x = [np.asarray(v, float) for v in x[:self.centers.shape[1]]]
        phi = None
        for c, r in zip(self.centers, self.radii):
            d = np.sqrt(sum((xi - ci) ** 2 for xi, ci in zip(x, c))) - r
            phi = d if phi is None else np.minimum(phi, d)
        return phi if self.fluid_inside else -phi


def Circle(center, radius):
    return Balls([list(center)], [radius])


Sphere = Circle


def Interval(center, radius):
    return Balls([[float(center)]], [radius])


class HalfSpace:
    """phi(x) = x[dim] - c ; ``-body`` flips."""
    kind = 1

    def __init__(self, dim, c, fluid_below=True):
        self.dim, self.c, self.fluid_inside = int(dim), float(c), bool(fluid_below)

    def __neg__(self):
        return HalfSpace(self.dim, self.c, not self.fluid_inside)

    def __call__(self, *x):
        phi = np.asarray(x[self.dim], float) - self.c
        return phi if self.fluid_inside else -phi


# ----------------------------------------------------------------------------------------------------------------
# Capacity  (src/capacity.jl:25-64, 81-123)
# ----------------------------------------------------------------------------------------------------------------
class Capacity:
    """``Capacity(body, mesh; method="VOFI", compute_centroids=true)``.

    Fields mirror src/capacity.jl:25-36; the reference stores diagonal sparse matrices, here each is the diagonal as a
    1-D array of the padded length (``V``, ``Γ``, ``A[d]``, ``B[d]``, ``W[d]``), ``C_ω``/``C_γ`` are (n, N) arrays and
    ``cell_types`` holds 1.0 / 0.0 / -1.0.  Arrays are fetched lazily from the device.
    """

    def __init__(self, body, mesh, method="VOFI", compute_centroids=True, _handle=None):
        self.mesh, self.body, self.N = mesh, body, mesh.N
        self._ctx = context()
        self._cache = {}
        self._has_cg = bool(compute_centroids)
        if _handle is not None:
            self._h = _handle
        else:
            if method not in ("VOFI", "ImplicitIntegration"):
                raise ValueError(f"unknown method {method}")
            if not isinstance(body, (Balls, HalfSpace)):
                raise TypeError("body must be a GPU-evaluable level set (Circle/Sphere/Balls/HalfSpace); "
                                "use Capacity.from_arrays for moments computed elsewhere")
            ls = L.LevelSetC()
            ls.kind = body.kind
            ls.fluid_inside = int(body.fluid_inside)
            if body.kind == 0:
                if body.centers.shape[1] != mesh.N:
                    raise ValueError("ball centres must have one coordinate per mesh dimension")
                cen = np.ascontiguousarray(body.centers, float)
                rad = np.ascontiguousarray(body.radii, float)
                ls.nballs, ls.centers, ls.radii = len(rad), _dp(cen), _dp(rad)
            else:
                ls.nballs, ls.hs_dim, ls.hs_c = 0, body.dim, body.c
            h = C.c_void_p()
            n = np.asarray(mesh.dims, np.int32)
            x0 = np.asarray(mesh.x0, float)
            Ls = np.asarray(mesh.domain_size, float)
            L.check(L.lib().pb200_capacity_create(self._ctx.h, mesh.N, n.ctypes.data_as(L.ip), _dp(x0), _dp(Ls), C.byref(ls),
                                                  int(compute_centroids), C.byref(h)), self._ctx.h)
            self._h = h
        k0, k1, nl = C.c_int(), C.c_int(), C.c_int64()
        L.check(L.lib().pb200_capacity_local(self._h, C.byref(k0), C.byref(k1), C.byref(nl)))
        self.k0, self.k1, self.nloc = k0.value, k1.value, nl.value

    @classmethod
    def from_arrays(cls, mesh, V, Γ, cell_types, A, B, W, C_ω, C_γ=None, body=None):
        """pb200_capacity_import: moments computed elsewhere (reference dumps, generic closures)."""
        ctx = context()
        N = mesh.N
        cat = lambda t: None if t is None else np.ascontiguousarray(np.concatenate([np.asarray(a, float).ravel() for a in t]))
        cols = lambda M: None if M is None else np.ascontiguousarray(np.asarray(M, float).T.ravel())
        Vc, Gc, Tc = (np.ascontiguousarray(np.asarray(a, float)) for a in (V, Γ, cell_types))
        Ac, Bc, Wc, Co, Cg = cat(A), cat(B), cat(W), cols(C_ω), cols(C_γ)
        h = C.c_void_p()
        n = np.asarray(mesh.dims, np.int32)
        x0 = np.asarray(mesh.x0, float)
        Ls = np.asarray(mesh.domain_size, float)
        L.check(L.lib().pb200_capacity_import(ctx.h, N, n.ctypes.data_as(L.ip), _dp(x0), _dp(Ls), _dp(Vc), _dp(Gc), _dp(Tc), _dp(Ac), _dp(Bc),
                                              _dp(Wc), _dp(Co), _dp(Cg), C.byref(h)), ctx.h)
        return cls(body, mesh, compute_centroids=C_γ is not None, _handle=h)

    def _fetch(self):
        if self._cache:
            return
        n, N = self.nloc, self.N
        V, G, T = np.empty(n), np.empty(n), np.empty(n)
        A, B, W, Co, Cg = (np.empty(N * n) for _ in range(5))
        L.check(L.lib().pb200_capacity_export(self._h, _dp(V), _dp(G), _dp(T), _dp(A), _dp(B), _dp(W), _dp(Co), _dp(Cg)), self._ctx.h)
        sp = lambda a: tuple(a[d * n:(d + 1) * n] for d in range(N))
        self._cache = dict(V=V, G=G, T=T, A=sp(A), B=sp(B), W=sp(W), Co=np.stack(sp(Co), axis=1), Cg=np.stack(sp(Cg), axis=1))

    def _get(self, k):
        self._fetch()
        return self._cache[k]

    V = property(lambda s: s._get("V"))
    Γ = property(lambda s: s._get("G"))
    Gamma = Γ
    cell_types = property(lambda s: s._get("T"))
    A = property(lambda s: s._get("A"))
    B = property(lambda s: s._get("B"))
    W = property(lambda s: s._get("W"))
    C_ω = property(lambda s: s._get("Co"))
    C_omega = C_ω
    C_γ = property(lambda s: s._get("Cg") if s._has_cg else np.empty((0, s.N)))
    C_gamma = C_γ

    def __del__(self):
        try:
            if getattr(self, "_h", None) and _ctx is not None:
                L.lib().pb200_capacity_destroy(self._h)
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------------------
# DiffusionOps, grad, div  (src/operators.jl:20-34, 49-55, 172-178)
# ----------------------------------------------------------------------------------------------------------------
class DiffusionOps:
    """``DiffusionOps(capacity)``: ``size`` = padded node counts; G and H are applied matrix-free on the device."""

    def __init__(self, capacity):
        self.capacity = capacity
        self._ctx = capacity._ctx
        self.size = tuple(d + 1 for d in capacity.mesh.dims)
        h = C.c_void_p()
        L.check(L.lib().pb200_ops_create(capacity._h, C.byref(h)), self._ctx.h)
        self._h = h

    @property
    def V(self):
        return self.capacity.V

    @property
    def Wdag(self):
        n, N = self.capacity.nloc, self.capacity.N
        out = np.empty(N * n)
        L.check(L.lib().pb200_ops_export_wdag(self._h, _dp(out)), self._ctx.h)
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) and _ctx is not None:
                L.lib().pb200_ops_destroy(self._h)
        except Exception:
            pass


class ConvectionOps(DiffusionOps):
    """``ConvectionOps(capacity, uₒ, uᵧ)`` (src/operators.jl:194-209): the diffusion operators plus the advective ones, C_d = D_p diag(S_m A_d uₒ_d) S_m
    and K_d = diag(S_p H' uᵧ), as coefficient arrays on the device (``pb200_ops_set_convection``).  ``uₒ``: N arrays of n bulk velocity components,
    ``uᵧ``: N n interface velocity components."""

    def __init__(self, capacity, uₒ, uᵧ):
        super().__init__(capacity)
        n, N = capacity.nloc, capacity.N
        uo = np.ascontiguousarray(np.concatenate([np.asarray(u, float).reshape(-1) for u in uₒ]))
        ug = np.ascontiguousarray(np.asarray(uᵧ, float).reshape(-1))
        if uo.shape != (N * n,) or ug.shape != (N * n,):
            raise ValueError("ConvectionOps: uₒ must hold N arrays of n values and uᵧ N n values")
        L.check(L.lib().pb200_ops_set_convection(self._h, _dp(uo), _dp(ug)), self._ctx.h)

    def coefficients(self):
        """(cf, kd): cf[d] = S_m A_d uₒ_d (C_d = D_p diag(cf[d]) S_m), kd = diag(0.5 Σ_d K_d) -- what the device rows read"""
        n, N = self.capacity.nloc, self.capacity.N
        cf, kd = np.empty(N * n), np.empty(n)
        L.check(L.lib().pb200_ops_export_convection(self._h, _dp(cf), _dp(kd)), self._ctx.h)
        return [cf[d * n:(d + 1) * n] for d in range(N)], kd


def grad(operator, p):
    """``∇(operator, p) = Wꜝ (G p_ω + H p_γ)`` (src/operators.jl:20-23)."""
    n, N = operator.capacity.nloc, operator.capacity.N
    p = np.ascontiguousarray(p, float)
    assert p.shape == (2 * n,)
    out = np.empty(N * n)
    L.check(L.lib().pb200_ops_grad(operator._h, _dp(p), _dp(out)), operator._ctx.h)
    return out


def div(operator, qω, qγ):
    """``∇₋(operator, qω, qγ) = -(Gᵀ + Hᵀ) qω + Hᵀ qγ`` (src/operators.jl:30-34)."""
    n = operator.capacity.nloc
    qω, qγ = np.ascontiguousarray(qω, float), np.ascontiguousarray(qγ, float)
    out = np.empty(n)
    L.check(L.lib().pb200_ops_div(operator._h, _dp(qω), _dp(qγ), _dp(out)), operator._ctx.h)
    return out


# ----------------------------------------------------------------------------------------------------------------
# boundary / phase value types  (src/boundary.jl:12-137, src/phase.jl:12-17)
# ----------------------------------------------------------------------------------------------------------------
class Dirichlet:
    def __init__(self, value): self.value = value


class Neumann:
    def __init__(self, value): self.value = value


class Robin:
    def __init__(self, α, β, value): self.α, self.β, self.value = α, β, value


class Periodic:
    pass


class ScalarJump:
    def __init__(self, α1, α2, value): self.α1, self.α2, self.value = α1, α2, value


class FluxJump:
    def __init__(self, β1, β2, value): self.β1, self.β2, self.value = β1, β2, value


class BorderConditions:
    def __init__(self, borders=None): self.borders = dict(borders or {})


class InterfaceConditions:
    def __init__(self, scalar, flux): self.scalar, self.flux = scalar, flux


class Phase:
    def __init__(self, capacity, operator, source, Diffusion_coeff):
        self.capacity, self.operator, self.source, self.Diffusion_coeff = capacity, operator, source, Diffusion_coeff


# ----------------------------------------------------------------------------------------------------------------
# closure evaluation, as the reference does it (src/solver.jl:230-323, 441-448)
# ----------------------------------------------------------------------------------------------------------------
def _coords3(Cm):
    n, N = Cm.shape
    return [Cm[:, d] for d in range(N)] + [np.zeros(n)] * (3 - N)


def _eval(f, cols, t, n):
    """-> (const, array): constants stay scalars so nothing is uploaded for them."""
    if not callable(f):
        return float(f), None
    if t is None:
        out = f(*cols)
    else:
        try:
            out = f(*cols, t)
        except TypeError:
            out = f(*cols)
    out = np.asarray(out, float)
    if out.ndim == 0:
        return float(out), None
    out = np.ascontiguousarray(np.broadcast_to(out, (n,)))
    if n and np.all(out == out[0]):
        return float(out[0]), None
    return 0.0, out


_SIDES = {"left": (0, 1, 0), "right": (1, 1, 1), "bottom": (2, 0, 0), "top": (3, 0, 1), "backward": (4, 2, 0), "forward": (5, 2, 1)}
_BCK = {Dirichlet: 1, Neumann: 2, Robin: 3, Periodic: 4}


def _border_values(mesh, key, cond, t):
    """eval_bc_value on every real cell of one side (positions = mesh.centers, src/mesh.jl:67)."""
    _, dim, hi = _SIDES[key]
    v = cond.value
    if not callable(v):
        return float(v), None
    ax = [mesh.centers[d] if d != dim else np.array([mesh.centers[d][-1 if hi else 0]]) for d in range(mesh.N)]
    grids = np.meshgrid(*ax, indexing="ij")
    pos = [g.ravel(order="F") for g in grids]          # other dims, x fastest
    if t is None:
        out = v(*pos)
    else:
        try:
            out = v(*pos, t)
        except TypeError:
            out = v(*pos)
    out = np.ascontiguousarray(np.broadcast_to(np.asarray(out, float), pos[0].shape))
    return 0.0, out


# ----------------------------------------------------------------------------------------------------------------
# Solver  (src/solver.jl:33-42) and the diffusion constructors / loops (src/solver/diffusion.jl)
# ----------------------------------------------------------------------------------------------------------------
class Solver:
    def __init__(self, time_type, phase_type, equation_type):
        self.time_type, self.phase_type, self.equation_type = time_type, phase_type, equation_type
        self.A = None          # never assembled (matrix-free)
        self.b = None
        self.x = None
        self.ch = []           # per-solve statistics (iterations, residual, device ms)
        self.states = []
        self._h = None
        self._ctx = context()
        self._first = None     # arguments of the constructor's step

    def __del__(self):
        try:
            if self._h and _ctx is not None:
                L.lib().pb200_solver_destroy(self._h)
        except Exception:
            pass


def _make_solver(s, phase1, phase2, bc_i, ic):
    d = L.SolverDesc()
    d.phase_type = 1 if phase2 is not None else 0
    d.time_type = 1 if s.time_type == Unsteady else 0
    d.ops1 = phase1.operator._h
    d.ops2 = phase2.operator._h if phase2 is not None else None
    keep = []
    for k, ph in ((1, phase1), (2, phase2)):
        if ph is None:
            continue
        Dc = ph.Diffusion_coeff
        cst, arr = _eval(Dc, _coords3(ph.capacity.C_ω) if callable(Dc) else None, None, ph.capacity.nloc)
        setattr(d, f"D{k}", cst)
        if arr is not None:
            keep.append(arr)
            setattr(d, f"D{k}_arr", _dp(arr))
    if phase2 is None:
        if isinstance(bc_i, Dirichlet):
            d.ifc_kind = 1
        elif isinstance(bc_i, Neumann):
            d.ifc_kind = 2
        elif isinstance(bc_i, Robin):
            d.ifc_kind, d.alpha, d.beta = 3, float(bc_i.α), float(bc_i.β)
        else:
            raise TypeError("interface condition must be Dirichlet, Neumann or Robin")
    else:
        d.alpha1, d.alpha2 = float(ic.scalar.α1), float(ic.scalar.α2)
        d.beta1, d.beta2 = float(ic.flux.β1), float(ic.flux.β2)
    h = C.c_void_p()
    L.check(L.lib().pb200_solver_create(s._ctx.h, C.byref(d), C.byref(h)), s._ctx.h)
    s._h = h
    s._phases = (phase1, phase2)


def _set_borders(s, mesh, bc_b, t):
    for key, cond in bc_b.borders.items():
        if key not in _SIDES:
            continue                                    # unknown keys (:front, :back) never match
        side, dim, _ = _SIDES[key]
        if dim >= mesh.N:
            continue
        kind = _BCK.get(type(cond), 0)
        cst, arr = (0.0, None)
        if kind == 1 or (kind == 2 and mesh.N == 1):                 # values matter for Dirichlet rows and for the 1-D Neumann row
            cst, arr = _border_values(mesh, key, cond, t)
            if kind == 2 and arr is not None:
                cst, arr = float(arr[0]), None
        L.check(L.lib().pb200_solver_set_border(s._h, side, kind, cst, _dp(arr)), s._ctx.h)


def _krylov_opts(method, kw):
    o = L.KrylovOpts()
    name = method if isinstance(method, str) else getattr(method, "__name__", "auto")
    name = name.lower()
    o.method = 1 if name == "cg" else 2 if name.startswith("bicgstab") else 0
    o.rtol = float(kw.get("reltol", kw.get("rtol", 1e-10)))
    o.atol = float(kw.get("abstol", kw.get("atol", 0.0)))
    o.maxit = int(kw.get("maxiter", 20000))
    o.warm_start = int(kw.get("warm_start", 1))
    o.check_every = int(kw.get("check_every", 4))
    o.path = {"auto": 0, "generic": 1, "folded": 2}[kw.get("path", "auto")]
    o.precond = {"default": 0, "mg": 1, "multigrid": 1}[kw.get("precond", "default")]
    return o


def _step(s, scheme, dt, t, bc_i, ic, opts, mono_border_t=True):
    """One ``solve_system!`` of the reference loop, on the device."""
    ph1, ph2 = s._phases
    si = L.StepIn()
    si.scheme = 1 if scheme == "CN" else 0
    si.dt = float(dt) if dt is not None else 0.0
    keep = []
    unsteady = s.time_type == Unsteady
    for k, ph in enumerate((ph1, ph2)):
        if ph is None:
            continue
        cols = _coords3(ph.capacity.C_ω) if callable(ph.source) else None
        times = (t, t + dt) if unsteady else (None,)
        for w, tt in enumerate(times):
            cst, arr = _eval(ph.source, cols, tt, ph.capacity.nloc)
            si.f_const[k][w] = cst
            if arr is not None:
                keep.append(arr)
                si.f_arr[k][w] = _dp(arr)
    if ph2 is None:
        if callable(bc_i.value):
            cols = _coords3(ph1.capacity.C_γ)
        else:
            cols = None
        times = (t, t + dt) if unsteady else (None,)
        for w, tt in enumerate(times):
            cst, arr = _eval(bc_i.value, cols, tt, ph1.capacity.nloc) if cols is not None else (float(bc_i.value), None)
            si.g_const[w] = cst
            if arr is not None:
                keep.append(arr)
                si.g_arr[w] = _dp(arr)
    else:
        for w, (bc, cap) in enumerate(((ic.scalar, ph1.capacity), (ic.flux, ph2.capacity))):
            if callable(bc.value):
                cst, arr = _eval(bc.value, _coords3(cap.C_γ), None, cap.nloc)      # no t (diffusion.jl:397)
            else:
                cst, arr = float(bc.value), None
            si.g_const[w] = cst
            if arr is not None:
                keep.append(arr)
                si.g_arr[w] = _dp(arr)
    st = L.StepStats()
    rc = L.lib().pb200_solver_step(s._h, C.byref(si), C.byref(opts), C.byref(st))
    L.check(rc, s._ctx.h, allow=(5,))
    n = ph1.capacity.nloc
    x = np.empty((4 if ph2 is not None else 2) * n)
    L.check(L.lib().pb200_solver_get_state(s._h, _dp(x)), s._ctx.h)
    s.x = x
    s.ch.append(dict(iters=st.iters, converged=bool(st.converged), rnorm=st.rnorm, bnorm=st.bnorm, solve_ms=st.solve_ms,
                     setup_ms=st.setup_ms, dof_bulk=st.dof_bulk, dof_ifc=st.dof_ifc, launches=st.launches,
                     apply_cells_uniform=st.apply_cells_uniform, apply_cells_general=st.apply_cells_general,
                     apply_cells_fast=st.apply_cells_fast))
    return st


def _set_state(s, x):
    x = np.ascontiguousarray(x, float)
    L.check(L.lib().pb200_solver_set_state(s._h, _dp(x)), s._ctx.h)


# ---- steady -----------------------------------------------------------------------------------------------------
def DiffusionSteadyMono(phase, bc_b, bc_i):
    """src/solver/diffusion.jl:14-28"""
    s = Solver(Steady, Monophasic, Diffusion)
    _make_solver(s, phase, None, bc_i, None)
    _set_borders(s, phase.capacity.mesh, bc_b, None)
    s._args = (bc_i, None)
    return s


def solve_DiffusionSteadyMono_(s, method="cg", algorithm=None, **kw):
    """src/solver/diffusion.jl:60-72"""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    _step(s, "BE", None, None, s._args[0], None, _krylov_opts(method, kw))
    return s


def DiffusionSteadyDiph(phase1, phase2, bc_b, ic):
    """src/solver/diffusion.jl:88-102"""
    s = Solver(Steady, Diphasic, Diffusion)
    _make_solver(s, phase1, phase2, None, ic)
    _set_borders(s, phase1.capacity.mesh, bc_b, None)
    s._args = (None, ic)
    return s


def solve_DiffusionSteadyDiph_(s, method="auto", algorithm=None, **kw):
    """src/solver/diffusion.jl:163-175"""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    _step(s, "BE", None, None, None, s._args[1], _krylov_opts(method, kw))
    return s


# ---- unsteady -----------------------------------------------------------------------------------------------------
def DiffusionUnsteadyMono(phase, bc_b, bc_i, Δt, Tᵢ, scheme):
    """src/solver/diffusion.jl:192-210: the constructor fixes the system of the FIRST solve (t = 0, ctor scheme)."""
    s = Solver(Unsteady, Monophasic, Diffusion)
    _make_solver(s, phase, None, bc_i, None)
    _set_state(s, Tᵢ)
    _set_borders(s, phase.capacity.mesh, bc_b, 0.0)
    s._first = ("CN" if scheme == "CN" else "BE", float(Δt), bc_i)
    return s


def solve_DiffusionUnsteadyMono_(s, phase, Δt, Tₑ, bc_b, bc, scheme, method="cg", algorithm=None, states_stride=1, **kw):
    """src/solver/diffusion.jl:268-301.  ``states_stride``: keep every k-th state on the host (the reference keeps all)."""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    opts = _krylov_opts(method, kw)
    sch0, dt0, bc0 = s._first
    t = 0.0
    _step(s, sch0, dt0, 0.0, bc0, None, opts)
    s.states.append(s.x)
    k = 0
    while t < Tₑ:
        t += Δt
        _set_borders(s, phase.capacity.mesh, bc_b, t)
        _step(s, scheme, Δt, t, bc, None, opts)
        k += 1
        if k % states_stride == 0:
            s.states.append(s.x)
    return s


def DiffusionUnsteadyDiph(phase1, phase2, bc_b, ic, Δt, Tᵢ, scheme):
    """src/solver/diffusion.jl:319-332"""
    s = Solver(Unsteady, Diphasic, Diffusion)
    _make_solver(s, phase1, phase2, None, ic)
    _set_state(s, Tᵢ)
    _set_borders(s, phase1.capacity.mesh, bc_b, None)      # BC_border_diph! is called without t (diffusion.jl:330)
    s._first = (scheme, float(Δt), ic)
    return s


def solve_DiffusionUnsteadyDiph_(s, phase1, phase2, Δt, Tₑ, bc_b, ic, scheme, method="auto", algorithm=None, states_stride=1, **kw):
    """src/solver/diffusion.jl:422-454"""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    opts = _krylov_opts(method, kw)
    sch0, dt0, ic0 = s._first
    t = 0.0
    _step(s, sch0, dt0, 0.0, None, ic0, opts)
    s.states.append(s.x)
    _set_borders(s, phase1.capacity.mesh, bc_b, None)
    k = 0
    while t < Tₑ:
        t += Δt
        _step(s, scheme, Δt, t, None, ic, opts)
        k += 1
        if k % states_stride == 0:
            s.states.append(s.x)
    return s


# ---- Darcy (src/solver/darcy.jl:1-89): the diffusion systems under another name + the velocity u = -∇p ---------------------------
# ---- advection-diffusion (src/solver/advectiondiffusion.jl) -- the same solver objects on ConvectionOps -------------------------------------
def AdvectionDiffusionSteadyMono(phase, bc_b, bc_i):
    """src/solver/advectiondiffusion.jl:12-28"""
    if not isinstance(phase.operator, ConvectionOps):
        raise TypeError("AdvectionDiffusionSteadyMono needs a phase built on ConvectionOps")
    s = Solver(Steady, Monophasic, DiffusionAdvection)
    _make_solver(s, phase, None, bc_i, None)
    _set_borders(s, phase.capacity.mesh, bc_b, None)
    s._args = (bc_i, None)
    return s


def solve_AdvectionDiffusionSteadyMono_(s, method="gmres", algorithm=None, **kw):
    """src/solver/advectiondiffusion.jl:65-71 (``gmres`` in the reference; the device solves the same rows with BiCGSTAB)"""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    _step(s, "BE", None, None, s._args[0], None, _krylov_opts(method, kw))
    return s


def AdvectionDiffusionUnsteadyMono(phase, bc_b, bc_i, Δt, Tᵢ, scheme):
    """src/solver/advectiondiffusion.jl:163-176: the constructor's system carries NO border rows (they enter in the loop, :273)."""
    if not isinstance(phase.operator, ConvectionOps):
        raise TypeError("AdvectionDiffusionUnsteadyMono needs a phase built on ConvectionOps")
    s = Solver(Unsteady, Monophasic, DiffusionAdvection)
    _make_solver(s, phase, None, bc_i, None)
    _set_state(s, Tᵢ)
    s._first = ("CN" if scheme == "CN" else "BE", float(Δt), bc_i)
    return s


def solve_AdvectionDiffusionUnsteadyMono_(s, phase, Δt, Tₑ, bc_b, bc, scheme, method="gmres", algorithm=None, states_stride=1, **kw):
    """src/solver/advectiondiffusion.jl:254-283 (the loop's RHS call with the diffusion coefficient the reference's text omits at :272)."""
    if s._h is None:
        raise RuntimeError("Solver is not initialized. Call a solver constructor first.")
    opts = _krylov_opts(method, kw)
    sch0, dt0, bc0 = s._first
    t = 0.0
    _step(s, sch0, dt0, 0.0, bc0, None, opts)
    s.states.append(s.x)
    k = 0
    while t < Tₑ:
        t += Δt
        _set_borders(s, phase.capacity.mesh, bc_b, t)
        _step(s, scheme, Δt, t, bc, None, opts)
        k += 1
        if k % states_stride == 0:
            s.states.append(s.x)
    return s


def DarcyFlow(phase, bc_b, bc_i):
    """src/solver/darcy.jl:1-15 -- the steady monophasic diffusion system (A_mono_stead_diff / b_mono_stead_diff + border rows)."""
    return DiffusionSteadyMono(phase, bc_b, bc_i)


def solve_DarcyFlow_(s, method="cg", algorithm=None, **kw):
    """src/solver/darcy.jl:17-24"""
    solve_DiffusionSteadyMono_(s, method=method, algorithm=algorithm, **kw)
    s.states.append(s.x)
    return s


def solve_darcy_velocity(solver, Fluide, state_i=1):
    """src/solver/darcy.jl:26-40: ``u = -∇(operator, p)`` of a stored state (1-based ``state_i`` as in the reference), with the
    pressures of cells that do not carry them set to NaN first: p_ω on empty cells, p_γ on empty and on full cells."""
    ct = Fluide.capacity.cell_types
    st = np.array(solver.states[state_i - 1], float)
    n = st.size // 2
    pω, pγ = st[:n].copy(), st[n:].copy()
    pω[ct == 0] = np.nan
    pγ[ct == 0] = np.nan
    pγ[ct == 1] = np.nan
    return -grad(Fluide.operator, np.concatenate([pω, pγ]))


def DarcyFlowUnsteady(phase, bc_b, bc_i, Δt, Tᵢ, scheme):
    """src/solver/darcy.jl:45-58"""
    return DiffusionUnsteadyMono(phase, bc_b, bc_i, Δt, Tᵢ, scheme)


def solve_DarcyFlowUnsteady_(s, phase, Δt, Tₑ, bc_b, bc_i, scheme, method="cg", algorithm=None, **kw):
    """src/solver/darcy.jl:60-89 (the time loop of solve_DiffusionUnsteadyMono! under another name)"""
    return solve_DiffusionUnsteadyMono_(s, phase, Δt, Tₑ, bc_b, bc_i, scheme, method=method, algorithm=algorithm, **kw)


# ---- check_convergence (src/convergence.jl:4-93) ------------------------------------------------------------------------------
def check_convergence(u_analytical, solver, capacity, p=2, relative=False, phase=0, verbose=False):
    """Volume-weighted error norms between ``u_analytical`` (evaluated on the host at ``capacity.C_ω``, as the reference does)
    and the bulk field of the solver's CURRENT device state, reduced on the device in one pass (no host copy of the state).
    Returns ``(u_ana, u_num, global_err, full_err, cut_err, empty_err)`` like the reference; ``u_num`` is ``None`` unless the
    solver's host copy ``solver.x`` exists (it is not downloaded for this).  ``phase`` (extension): 1 = second bulk field of a
    diphasic solver (the reference handles monophasic solvers only: ``solver.x[1:end÷2]``, src/convergence.jl:73)."""
    Cω = capacity.C_ω
    N = capacity.N
    u_ana = np.ascontiguousarray(np.asarray(u_analytical(*[Cω[:, d] for d in range(N)]), float) * np.ones(Cω.shape[0]))
    out = (C.c_double * 4)()
    L.check(L.lib().pb200_solver_error_norms(solver._h, int(phase), _dp(u_ana), float(p), 1 if relative else 0,
                                             C.cast(out, L.dp)), solver._ctx.h)
    g, fu, cu, em = (float(v) for v in out)
    if verbose:
        print(f"All cells L{p} norm        = {g}\nFull cells L{p} norm   = {fu}\nCut cells L{p} norm    = {cu}\nEmpty cells L{p} norm  = {em}")
    u_num = None
    if solver.x is not None:
        n = Cω.shape[0]
        u_num = np.asarray(solver.x)[2 * phase * n:(2 * phase + 1) * n]
    return u_ana, u_num, g, fu, cu, em
