#!/usr/bin/env python
"""Generates tests/golden/geometry_golden.json: capacities of two tiny meshes evaluated INDEPENDENTLY of oracle/ and of the
CUDA kernels, with nested chord integrals split at every kink: mpmath (30 digits, tanh-sinh) for the 1-D / 2-D cases, SciPy QUADPACK in
double precision for the 3-D cases.  Run:  python tests/golden/make_geometry_golden.py   (about 20 minutes on one core).

Meaning of each array: SURVEY.md section 8a row a3 (/root/reference/src/capacity.jl:81-123 and, for the definitions,
src/front_tracking.jl:814-1427): V fluid volume of cell [nodes_i, nodes_i+1]^N, C_omega its barycentre (cell centre when the
cell is empty), Gamma interface measure, C_gamma its centroid, A_d wetted measure of the lower face, B_d wetted section
through the barycentre, W_d wetted volume of the box between the barycentres of cells i-1 and i (i = 1..n_d-1, 0-based).
Padded layout n = prod(n_i + 1), x fastest, pad = 0; geometry grid = x0 + (j + 1/2) h (src/mesh.jl:50).
"""
import itertools
import json
import os

import math

import mpmath
from scipy.integrate import quad as sp_quad

mpmath.mp.dps = 30
# Two arithmetic back-ends behind the same code.  "mp": 30-digit mpmath, tanh-sinh quadrature -- used for the 1-D and 2-D cases.
# "float": IEEE double with SciPy's QUADPACK (adaptive Gauss-Kronrod 21, split at the same kinks, epsabs 1e-16) -- used for the 3-D
# cases, where three nested 30-digit quadratures take hours; its accuracy (~1e-13 of a cell) is what bounds the 3-D tolerances.
mpf, sqrt, acos, asin, cos, sin, pi = mpmath.mpf, mpmath.sqrt, mpmath.acos, mpmath.asin, mpmath.cos, mpmath.sin, mpmath.pi
BACKEND = ["mp"]


def set_backend(name):
    global mpf, sqrt, acos, asin, cos, sin, pi
    BACKEND[0] = name
    if name == "mp":
        mpf, sqrt, acos, asin, cos, sin, pi = mpmath.mpf, mpmath.sqrt, mpmath.acos, mpmath.asin, mpmath.cos, mpmath.sin, mpmath.pi
    else:
        mpf, sqrt, acos, asin, cos, sin, pi = float, math.sqrt, math.acos, math.asin, math.cos, math.sin, math.pi


def quad(f, pts):
    if BACKEND[0] == "mp":
        return mpmath.quad(f, pts)
    tot = 0.0
    for a, b in zip(pts[:-1], pts[1:]):
        if b - a > 1e-15 * (abs(a) + abs(b) + 1e-300):
            tot += sp_quad(f, a, b, epsabs=1e-16, epsrel=1e-14, limit=400)[0]
    return tot


def chord(c, r2, lo, hi, mid):
    """length and first moment (about mid) of [c - r, c + r] /\\ [lo, hi]"""
    if r2 <= 0:
        return mpf(0), mpf(0)
    r = sqrt(r2)
    a, b = max(lo, c - r), min(hi, c + r)
    if b <= a:
        return mpf(0), mpf(0)
    return b - a, ((b - mid) ** 2 - (a - mid) ** 2) / 2


def breakpoints(c0, R2, others, a, b):
    """abscissae along axis 0 where the clipped section changes its analytic form"""
    pts = {a, b}
    cand = [mpf(0)]
    # every subset of the remaining axes, every lo/hi choice (edges / corners of the cross-section box)
    dims = list(range(len(others)))
    for k in range(1, len(dims) + 1):
        for sub in itertools.combinations(dims, k):
            for ch in itertools.product((0, 1), repeat=k):
                cand.append(sum((others[d][2 + s] - others[d][0]) ** 2 for d, s in zip(sub, ch)))
    for d2 in cand:
        if d2 < R2:
            s = sqrt(R2 - d2)
            for x in (c0 - s, c0 + s):
                if a < x < b:
                    pts.add(x)
    return sorted(pts)


def ball_box(c, R2, lo, hi):
    """measure and first moments (about the box centre) of ball /\\ box, any dimension, nested quadrature, axis 0 outermost"""
    m = len(c)
    mid = [(l + h) / 2 for l, h in zip(lo, hi)]
    if m == 0:
        return [mpf(1) if R2 > 0 else mpf(0)]
    if R2 <= 0:
        return [mpf(0)] * (m + 1)
    if m == 1:
        ln, mo = chord(c[0], R2, lo[0], hi[0], mid[0])
        return [ln, mo]
    R = sqrt(R2)
    a, b = max(lo[0], c[0] - R), min(hi[0], c[0] + R)
    if b <= a:
        return [mpf(0)] * (m + 1)
    others = [(c[d], None, lo[d], hi[d]) for d in range(1, m)]
    pts = breakpoints(c[0], R2, others, a, b)
    out = []
    for q in range(m + 1):
        def f(x, q=q):
            sub = ball_box(c[1:], R2 - (x - c[0]) ** 2, lo[1:], hi[1:])
            if q == 0:
                return sub[0]
            if q == 1:
                return (x - mid[0]) * sub[0]
            return sub[q - 1]
        out.append(quad(f, pts))
    return out


def circle_rect_arcs(c, rho, lo, hi):
    """arc length and first moments (about the rectangle centre) of the circle inside the rectangle, exact angles"""
    mid = [(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2]
    ang = []
    for side in (0, 1):
        u = ((hi[0] if side else lo[0]) - c[0]) / rho
        if abs(u) < 1:
            a = acos(u)
            ang += [a, 2 * pi - a]
        v = ((hi[1] if side else lo[1]) - c[1]) / rho
        if abs(v) < 1:
            a = asin(v)
            ang += [a if a >= 0 else a + 2 * pi, pi - a]
    inside = lambda x, y: lo[0] <= x <= hi[0] and lo[1] <= y <= hi[1]
    if not ang:
        if inside(c[0] + rho, c[1]) and inside(c[0] - rho, c[1]) and inside(c[0], c[1] + rho) and inside(c[0], c[1] - rho):
            L = 2 * pi * rho
            return [L, L * (c[0] - mid[0]), L * (c[1] - mid[1])]
        return [mpf(0)] * 3
    ang = sorted(ang)
    ang.append(ang[0] + 2 * pi)
    out = [mpf(0)] * 3
    for a0, a1 in zip(ang[:-1], ang[1:]):
        if a1 <= a0:
            continue
        am = (a0 + a1) / 2
        if inside(c[0] + rho * cos(am), c[1] + rho * sin(am)):
            d = a1 - a0
            out[0] += rho * d
            out[1] += rho * ((c[0] - mid[0]) * d + rho * (sin(a1) - sin(a0)))
            out[2] += rho * ((c[1] - mid[1]) * d - rho * (cos(a1) - cos(a0)))
    return out


def sphere_box(c, R, lo, hi):
    """interface measure and first moments about the box centre (N = 1: points, 2: arcs, 3: hat-box form S = R int dphi dx)"""
    m = len(c)
    mid = [(l + h) / 2 for l, h in zip(lo, hi)]
    if m == 1:
        out = [mpf(0), mpf(0)]
        for p in (c[0] - R, c[0] + R):
            if lo[0] <= p < hi[0]:
                out[0] += 1
                out[1] += p - mid[0]
        return out
    if m == 2:
        return circle_rect_arcs(c, R, lo, hi)
    a, b = max(lo[0], c[0] - R), min(hi[0], c[0] + R)
    if b <= a:
        return [mpf(0)] * 4
    others = [(c[d], None, lo[d], hi[d]) for d in range(1, 3)]
    pts = breakpoints(c[0], R * R, others, a, b)
    out = []
    for q in range(4):
        def f(x, q=q):
            r2 = R * R - (x - c[0]) ** 2
            if r2 <= 0:
                return mpf(0)
            rho = sqrt(r2)
            sub = circle_rect_arcs(c[1:], rho, lo[1:], hi[1:])
            fac = R / rho
            if q == 0:
                return fac * sub[0]
            if q == 1:
                return fac * (x - mid[0]) * sub[0]
            return fac * sub[q - 1]
        out.append(quad(f, pts))
    return out


def capacity(nc, L, x0, center, R, inside):
    N = len(nc)
    set_backend("float" if N == 3 else "mp")
    h = [mpf(L[d]) / nc[d] for d in range(N)]
    nodes = [[mpf(x0[d]) + (mpf(j) + mpf(1) / 2) * h[d] for j in range(nc[d] + 2)] for d in range(N)]
    pd = [v + 1 for v in nc]
    n = 1
    for v in pd:
        n *= v
    c = [mpf(v) for v in center]
    R = mpf(R)
    lin = lambda ix: sum(ix[d] * (1 if d == 0 else (pd[0] if d == 1 else pd[0] * pd[1])) for d in range(N))

    def fluid(dims, lo, hi, fixd=None, fixv=None):
        """fluid measure + first moments about the box centre of a box spanning `dims`, optional fixed coordinate"""
        R2 = R * R
        if fixd is not None:
            R2 = R2 - (fixv - c[fixd]) ** 2
        cc = [c[d] for d in dims]
        mom = ball_box(cc, R2, lo, hi) if dims else [mpf(1) if R2 > 0 else mpf(0)]
        full = mpf(1)
        for l, hh in zip(lo, hi):
            full *= hh - l
        if not inside:
            mom = [full - mom[0]] + [-v for v in mom[1:]]
        return mom, full

    zeros = lambda k=1: [[mpf(0)] * n for _ in range(k)]
    V, Gam = zeros()[0], zeros()[0]
    ct = [0.0] * n
    A, B, W, Co, Cg = zeros(N), zeros(N), zeros(N), zeros(N), zeros(N)
    cells = list(itertools.product(*[range(v) for v in reversed(nc)]))
    for rix in cells:
        ix = tuple(reversed(rix))
        idx = lin(ix)
        lo = [nodes[d][ix[d]] for d in range(N)]
        hi = [nodes[d][ix[d] + 1] for d in range(N)]
        mid = [(l + hh) / 2 for l, hh in zip(lo, hi)]
        mom, full = fluid(list(range(N)), lo, hi)
        sur = sphere_box(c, R, lo, hi)
        eps = mpf(10) ** -25
        V[idx] = mom[0]
        if sur[0] > eps or (eps * full < mom[0] < full * (1 - eps)):
            ct[idx] = -1.0
        else:
            ct[idx] = 1.0 if mom[0] > full / 2 else 0.0
            V[idx] = full if ct[idx] == 1.0 else mpf(0)
        for d in range(N):
            Co[d][idx] = mid[d] + (mom[1 + d] / mom[0] if (ct[idx] == -1.0 and mom[0] > 0) else 0)
        if ct[idx] == -1.0:
            Gam[idx] = sur[0]
            if sur[0] > 0:
                for d in range(N):
                    Cg[d][idx] = mid[d] + sur[1 + d] / sur[0]
    for d in range(N):
        od = [e for e in range(N) if e != d]
        pcells = list(itertools.product(*[range(v) for v in reversed(pd)]))
        for rix in pcells:
            ix = tuple(reversed(rix))
            if any(ix[e] >= nc[e] for e in od):
                continue
            idx = lin(ix)
            lo = [nodes[e][ix[e]] for e in od]
            hi = [nodes[e][ix[e] + 1] for e in od]
            mom, face = fluid(od, lo, hi, d, nodes[d][ix[d]])
            A[d][idx] = mom[0]
            if ix[d] < nc[d]:
                if ct[idx] == 1.0:
                    B[d][idx] = face
                elif ct[idx] == -1.0:
                    B[d][idx] = fluid(od, lo, hi, d, Co[d][idx])[0][0]
                if ix[d] >= 1:
                    blo = [nodes[e][ix[e]] for e in range(N)]
                    bhi = [nodes[e][ix[e] + 1] for e in range(N)]
                    str_d = 1 if d == 0 else (pd[0] if d == 1 else pd[0] * pd[1])
                    blo[d], bhi[d] = Co[d][idx - str_d], Co[d][idx]
                    W[d][idx] = fluid(list(range(N)), blo, bhi)[0][0]
    f = lambda a: [float(v) for v in a]
    return dict(n=list(nc), L=list(L), x0=list(x0), center=list(center), radius=float(R), fluid_inside=bool(inside),
                V=f(V), Gamma=f(Gam), cell_types=ct, A=[f(a) for a in A], B=[f(a) for a in B], W=[f(a) for a in W],
                C_omega=[f(a) for a in Co], C_gamma=[f(a) for a in Cg])


if __name__ == "__main__":
    cases = {
        "circle_5x4_inside": capacity((5, 4), (4.0, 3.0), (0.0, 0.0), (1.9, 1.45), 1.05, True),
        "circle_5x4_outside": capacity((5, 4), (4.0, 3.0), (0.0, 0.0), (1.9, 1.45), 1.05, False),
        "interval_6_inside": capacity((6,), (4.0,), (0.0,), (2.03,), 0.95, True),   # end points off the grid nodes
        "sphere_3x3x3_inside": capacity((3, 3, 3), (4.0, 4.0, 4.0), (0.0, 0.0, 0.0), (1.9, 2.1, 2.05), 1.2, True),
        "sphere_3x3x3_outside": capacity((3, 3, 3), (4.0, 4.0, 4.0), (0.0, 0.0, 0.0), (1.9, 2.1, 2.05), 1.2, False),
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geometry_golden.json")
    with open(out, "w") as fh:
        json.dump(dict(generator="tests/golden/make_geometry_golden.py (1-D/2-D: mpmath dps=30; 3-D: nested QUADPACK in double precision)", cases=cases), fh)
    print("wrote", out)
