"""The C-ABI library loads and exports every symbol include/penguin_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "penguin_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    from penguin_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    so = ctypes.CDLL(_lib.SO)
    for nm in names:
        assert hasattr(so, nm), f"{nm} declared in include/penguin_b200.h but not exported by libpenguin_b200.so"
    assert sorted(_lib.SYMBOLS) == names, "ctypes binding table and header disagree"
    _lib.lib()   # resolves argtypes / restype of all of them


def test_struct_layouts_match_the_header(tmp_path):
    # sizeof() of every struct as gcc sees the header vs the ctypes mirror: a drifting mirror would corrupt calls silently
    import subprocess
    from penguin_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "penguin_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(pb200_levelset), '
                   'sizeof(pb200_solver_desc), sizeof(pb200_step_in), sizeof(pb200_krylov_opts), sizeof(pb200_step_stats)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    mirror = [ctypes.sizeof(c) for c in (_lib.LevelSetC, _lib.SolverDesc, _lib.StepIn, _lib.KrylovOpts, _lib.StepStats)]
    assert sizes == mirror


def test_no_cpu_fallback():
    # without a CUDA device the library must fail loudly (PB200_ENODEV), never compute on the host
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from penguin_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib().pb200_init(ctypes.byref(h), 0)
    assert rc == 2
    assert b"no CPU fallback" in _lib.lib().pb200_last_error(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "penguin.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "libgeom_oracle" not in src, f
                assert not re.search(r'#include\s+"[^"]*oracle', src), f


def test_polynomial_preconditioner_coefficients():
    """pb200_poly_coefs (host-only) against a NumPy Chebyshev iteration: z_{m+1} = q_m(A) r for A z = r, z_0 = 0 on [lo, hi]."""
    import ctypes
    import numpy as np
    from penguin_b200 import _lib
    lo, hi = 1.0 / 3.0, 1.95
    out = (ctypes.c_double * 6)()
    assert _lib.lib().pb200_poly_coefs(lo, hi, ctypes.cast(out, _lib.dp)) == 0
    c = list(out)
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.standard_normal((40, 40)))
    A = Q @ np.diag(np.linspace(0.2, 1.9, 40)) @ Q.T
    r = rng.standard_normal(40)

    def cheb(m):                      # residual form of the Chebyshev iteration, m matvecs
        theta, delta = (hi + lo) / 2, (hi - lo) / 2
        sigma = theta / delta
        rho = 1 / sigma
        z, res, d = np.zeros(40), r.copy(), r / theta
        for _ in range(m):
            z = z + d
            res = res - A @ d
            rho_n = 1 / (2 * sigma - rho)
            d = rho_n * rho * d + 2 * rho_n / delta * res
            rho = rho_n
        return z + d
    z2 = c[0] * r + c[2] * (A @ r)
    z3 = c[3] * r + c[4] * z2 + c[5] * (A @ z2)
    assert c[1] == 0.0
    assert np.allclose(z2, cheb(1), rtol=1e-12, atol=1e-13) and np.allclose(z3, cheb(2), rtol=1e-12, atol=1e-13)
    # q(A) is positive definite on the whole spectrum, also below lo (the interface modes) and up to hi
    lam = np.linspace(0.01, hi, 200)
    assert np.all(c[0] + c[2] * lam > 0)
    assert _lib.lib().pb200_poly_coefs(1.0, 0.5, ctypes.cast(out, _lib.dp)) != 0


def test_krylov_opts_mirror_maps_the_preconditioner():
    # pb200_krylov_opts.precond (include/penguin_b200.h): the Python mirror's `precond="mg"` and the Julia shim's `precond=:mg` select PB200_PRECOND_MG
    from penguin_b200 import api
    hdr = open(os.path.join(ROOT, "include", "penguin_b200.h")).read()
    assert re.search(r"#define\s+PB200_PRECOND_DEFAULT\s+0\b", hdr) and re.search(r"#define\s+PB200_PRECOND_MG\s+1\b", hdr)
    o = api._krylov_opts("cg", dict(precond="mg", reltol=1e-9))
    assert (o.precond, o.method, o.rtol) == (1, 1, 1e-9)
    assert api._krylov_opts("cg", {}).precond == 0
    with pytest.raises(KeyError):
        api._krylov_opts("cg", dict(precond="ilu"))
    jl = open(os.path.join(ROOT, "julia", "b200.jl")).read()
    fields = re.search(r"struct pb200_krylov_opts\s*\n\s*(.*?)\nend", jl, flags=re.S).group(1)
    names_jl = re.findall(r"(\w+)::C", fields)
    block = re.search(r"typedef struct \{([^}]*)\}\s*pb200_krylov_opts;", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S), flags=re.S).group(1)
    names_h = re.findall(r"\b(?:int|double)\s+(\w+)\s*;", block)
    assert names_jl == names_h == [f[0] for f in api.L.KrylovOpts._fields_]
