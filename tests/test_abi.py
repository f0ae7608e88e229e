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
