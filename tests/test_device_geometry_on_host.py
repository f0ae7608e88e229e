"""The geometry primitives of the CUDA capacity kernels (penguin.jl_b200/csrc/geometry.cuh: exact disc /\\ rectangle, adaptive
Gauss-Kronrod ball /\\ box) are __host__ __device__: tests/host_harness/geom_primitives.cu compiles THAT source for the CPU and runs it
against the C oracle on random configurations -- a check of the shipped arithmetic that needs no GPU.  Disagreements beyond 1e-12 are
arbitrated with mpmath (which side is off), closed forms pin the degenerate alignments (ball centre on a face / edge / corner)."""
import os
import shutil
import subprocess

import pytest

from oracle import geom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def harness_output(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    so = geom.build()
    exe = str(tmp_path_factory.mktemp("hh") / "geom_primitives")
    src = os.path.join(ROOT, "tests", "host_harness", "geom_primitives.cu")
    subprocess.check_call([NVCC if os.path.exists(NVCC) else "nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", src, "-o", exe, so,
                           "-Xlinker", "-rpath=" + os.path.dirname(so)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return subprocess.run([exe, "200000", "3000"], check=True, capture_output=True, text=True).stdout.splitlines()


def _exact_area(cx, cy, R, hx, hy):
    import mpmath as mp
    mp.mp.dps = 40
    cx, cy, R, hx, hy = [mp.mpf(v) for v in (cx, cy, R, hx, hy)]

    def chord(x):
        d2 = R * R - (x - cx) ** 2
        if d2 <= 0:
            return mp.mpf(0)
        s = mp.sqrt(d2)
        return max(min(hy, cy + s) - max(-hy, cy - s), mp.mpf(0))
    a, b = max(-hx, cx - R), min(hx, cx + R)
    if b <= a:
        return mp.mpf(0)
    pts = [a, b]
    for yy in (-hy, hy):
        d2 = R * R - (yy - cy) ** 2
        if d2 > 0:
            s = mp.sqrt(d2)
            pts += [xx for xx in (cx - s, cx + s) if a < xx < b]
    return mp.quad(chord, sorted(pts))


def test_disc_rect_vs_oracle_200k_random(harness_output):
    summ = [ln.split() for ln in harness_output if ln.startswith("SUMMARY2D")][0]
    n, bad = int(summ[1]), int(summ[2])
    assert n == 200000 and bad <= 20, f"{bad} of {n} disc/rectangle configurations differ from the oracle by more than 1e-12"
    pytest.importorskip("mpmath")
    for ln in harness_output:
        if not ln.startswith("BAD2D"):
            continue
        _, it, cx, cy, R, hx, hy, dev, orc, _, _ = ln.split()
        cx, cy, R, hx, hy, dev, orc = map(float, (cx, cy, R, hx, hy, dev, orc))
        ex = float(_exact_area(cx, cy, R, hx, hy))
        cell = 4 * hx * hy
        # the exact formula of the device code is the accurate side; the oracle's quadrature is what drifts to ~1e-11 near tangency
        assert abs(dev - ex) <= 2e-14 * cell, (it, dev, ex)
        assert abs(orc - ex) <= 1e-9 * cell, (it, orc, ex)


def test_ball_box_vs_oracle_and_closed_forms(harness_output):
    summ = [ln.split() for ln in harness_output if ln.startswith("SUMMARY3D")][0]
    n, bad, wv, ws = int(summ[1]), int(summ[2]), float(summ[3]), float(summ[4])
    assert n == 3000 and bad == 0 and wv < 2e-12 and ws < 5e-12
    closed = [ln.split() for ln in harness_output if ln.startswith("CLOSED")]
    assert len(closed) == 4
    for _, k, v, vex, s, sex in closed:
        # ball of radius 0.37 with k coordinates of its centre ON the box boundary: 1 / 2^k of the ball, volume and surface
        # (k = 1 was off by 2e-10 with the asin-based arc angle: every z-section has its centre on an edge line)
        assert abs(float(v) - float(vex)) <= 1e-14 and abs(float(s) - float(sex)) <= 1e-13, (k, v, vex, s, sex)


def test_tangency_at_a_grid_node(harness_output):
    # test/solver/darcy_test.jl:10 -- circle (0.5, 0.5) r 0.5 on the h = 0.1 grid touches x = 0 exactly at a cell corner
    mp = pytest.importorskip("mpmath")
    rows = [ln.split() for ln in harness_output if ln.startswith("TANGENT")]
    assert len(rows) == 8
    for _, i, j, lox, hix, loy, hiy, area, arc in rows:
        lox, hix, loy, hiy, area, arc = map(float, (lox, hix, loy, hiy, area, arc))
        mx, my = 0.5 * (lox + hix), 0.5 * (loy + hiy)
        ex = float(_exact_area(0.5 - mx, 0.5 - my, 0.5, 0.5 * (hix - lox), 0.5 * (hiy - loy)))
        assert abs(area - ex) <= 1e-15, (i, j, area, ex)
        if i == "1":
            assert area == pytest.approx(0.01, abs=1e-17) and arc == 0.0          # second column: full cells, no interface
        else:
            assert 0.0 < area < 0.01 and arc > 0.1                                  # first column: cut cells
