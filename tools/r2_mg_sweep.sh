#!/bin/bash
# Multigrid preconditioner of configs[4] (csrc/mg.cuh): parity tests, then the knobs of the V-cycle at 256^3 (called through gpurun).
O=gpurun_out
timeout 150 python -m pytest tests/test_gpu_zz_poisson3d.py -x -q > $O/r2_mg_test.log 2>&1; tail -2 $O/r2_mg_test.log
run() {   # tag, env...
    tag=$1; shift
    env "$@" timeout 100 python tools/run_poisson3d.py --nx ${NX:-256} --precond mg --repeat 2 > $O/r2_mg_$tag.json 2> $O/r2_mg_$tag.err
    python - "$tag" $O/r2_mg_$tag.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:28s} it {d['iterations']:4d}  loop {d['krylov_loop_ms']:8.1f} ms  {d['ms_per_iteration']:6.3f} ms/it  launches {d['launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default A=1
run nograph PB200_NO_GRAPH=1
run lev4 PB200_MG_LEVELS=4
run lev5 PB200_MG_LEVELS=5
run lev5_sw30 PB200_MG_LEVELS=5 PB200_MG_SWEEPS=30
run lev4_sw40_ac150 PB200_MG_LEVELS=4 PB200_MG_SWEEPS=40 PB200_MG_ALPHAC=150
run alpha4 PB200_MG_ALPHA=4
run alpha16 PB200_MG_ALPHA=16
run deg1_alpha4 PB200_MG_DEG=1 PB200_MG_ALPHA=4
run sw4 PB200_MG_SWEEPS=4
NX=384 run n384_default A=1
