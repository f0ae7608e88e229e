"""ctypes front-end of oracle/geom_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

``capacity(mesh, levelset)`` returns an ``oracle.penguin_oracle.Capacity`` whose arrays follow the layout of
/root/reference/src/capacity.jl:81-123 (padded n = prod(n_i+1), x fastest).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import penguin_oracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgeom_oracle.so")
_SRC = os.path.join(_HERE, "geom_oracle.c")
_lib = None


def build(force=False):
    """gcc build of the C restatement (called by __graft_entry__.build() and lazily by tests)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", _SRC, "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.pgo_capacity.argtypes = [ctypes.c_int, ip, dp, dp, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_double] + [dp] * 8
        _lib.pgo_capacity.restype = ctypes.c_int
        _lib.pgo_ball_box.argtypes = [ctypes.c_int, dp, ctypes.c_double, dp, dp, dp]
        _lib.pgo_sphere_box.argtypes = [ctypes.c_int, dp, ctypes.c_double, dp, dp, dp]
    return _lib


class LevelSet:
    """GPU-evaluable level-set descriptor shared by the oracle and the product's host API.

    kind 'balls': phi(x) = min_k(|x - c_k| - r_k) for disjoint balls; kind 'halfspace': phi = x[dim] - c.
    ``fluid_inside=True`` means fluid = {phi < 0}; False is the sign flip (the reference's ``-(...)`` bodies).
    """

    def __init__(self, kind, centers=None, radii=None, fluid_inside=True, dim=0, c=0.0):
        self.kind = kind
        self.centers = None if centers is None else np.atleast_2d(np.asarray(centers, float))
        self.radii = None if radii is None else np.atleast_1d(np.asarray(radii, float))
        self.fluid_inside = bool(fluid_inside)
        self.dim, self.c = int(dim), float(c)

    @staticmethod
    def ball(center, radius, fluid_inside=True):
        return LevelSet("balls", [list(np.atleast_1d(center))], [radius], fluid_inside)

    @staticmethod
    def balls(centers, radii, fluid_inside=True):
        return LevelSet("balls", centers, radii, fluid_inside)

    @staticmethod
    def halfspace(dim, c, fluid_below=True):
        return LevelSet("halfspace", None, None, fluid_below, dim, c)

    def flipped(self):
        return LevelSet(self.kind, self.centers, self.radii, not self.fluid_inside, self.dim, self.c)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def capacity(mesh: po.Mesh, ls: LevelSet, compute_centroids=True) -> po.Capacity:
    N, n = mesh.N, mesh.n
    nc = np.asarray(mesh.dims, np.int32)
    x0 = np.asarray(mesh.x0, float)
    L = np.asarray(mesh.L, float)
    V, G, ct = np.zeros(n), np.zeros(n), np.zeros(n)
    A, B, W, Co = np.zeros(N * n), np.zeros(N * n), np.zeros(N * n), np.zeros(N * n)
    Cg = np.zeros(N * n)
    if ls.kind == "balls":
        cen = np.ascontiguousarray(ls.centers, float)
        rad = np.ascontiguousarray(ls.radii, float)
        assert cen.shape == (len(rad), N)
        kind, nb = 0, len(rad)
    else:
        cen, rad, kind, nb = np.zeros(1), np.zeros(1), 1, 0
    rc = lib().pgo_capacity(N, nc.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(x0), _dp(L), kind, nb, _dp(cen), _dp(rad),
                            int(ls.fluid_inside), ls.dim, ls.c, _dp(V), _dp(G), _dp(ct), _dp(A), _dp(B), _dp(W), _dp(Co),
                            _dp(Cg) if compute_centroids else None)
    if rc != 0:
        raise RuntimeError(f"pgo_capacity failed: {rc}")
    sp = lambda a: tuple(a[d * n:(d + 1) * n].copy() for d in range(N))
    return po.Capacity(mesh, V, G, ct, sp(A), sp(B), sp(W), np.stack(sp(Co), axis=1),
                       np.stack(sp(Cg), axis=1) if compute_centroids else None)


def ball_box(c, R, lo, hi):
    c, lo, hi = (np.ascontiguousarray(v, float) for v in (c, lo, hi))
    out = np.zeros(len(c) + 1)
    lib().pgo_ball_box(len(c), _dp(c), float(R), _dp(lo), _dp(hi), _dp(out))
    return out


def sphere_box(c, R, lo, hi):
    c, lo, hi = (np.ascontiguousarray(v, float) for v in (c, lo, hi))
    out = np.zeros(len(c) + 1)
    lib().pgo_sphere_box(len(c), _dp(c), float(R), _dp(lo), _dp(hi), _dp(out))
    return out
