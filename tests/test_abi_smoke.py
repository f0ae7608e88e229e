"""tests/abi_smoke.c: the C ABI walked from plain C (gcc + dlopen) in the order of the Julia shim.  Without a GPU the program must load the
library, resolve every symbol it uses and stop at pb200_init with PB200_ENODEV (exit code 77); with one it runs the diphasic problem and its
final state must equal, bit for bit, what the Python mirror of the shim computes through the same calls."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"), "-ldl", "-lm", "-o", exe])
    return exe


def test_c_program_builds_loads_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: see the gpu test")
    from penguin_b200 import _lib
    out = subprocess.run([_build(tmp_path), _lib.SO], capture_output=True, text=True, timeout=120)
    assert out.returncode == 77, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_c_walk_equals_the_python_mirror(tmp_path):
    import penguin_b200 as pb
    from penguin_b200 import _lib
    state_file = str(tmp_path / "state.bin")
    out = subprocess.run([_build(tmp_path), _lib.SO, state_file], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ABI_SMOKE_OK" in out.stdout, out.stdout + out.stderr
    xc = np.fromfile(state_file)
    pb.init()
    nx = 64
    mesh = pb.Mesh((nx, nx), (8.0, 8.0))
    body = pb.Circle((4.0, 4.0), 2.0)
    c1, c2 = pb.Capacity(body, mesh), pb.Capacity(-body, mesh)
    p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
    n = c1.nloc
    dt = 0.5 * (8.0 / nx) ** 2
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    bc = pb.BorderConditions({"left": pb.Dirichlet(0.0)})
    s = pb.DiffusionUnsteadyDiph(p1, p2, bc, ic, dt, np.concatenate([np.ones(2 * n), np.zeros(2 * n)]), "BE")
    pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 1.5 * dt, bc, ic, "BE", reltol=1e-12, warm_start=1, check_every=4)
    assert len(s.states) == 3
    assert np.array_equal(s.states[-1], xc)
