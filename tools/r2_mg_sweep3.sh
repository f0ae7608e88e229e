#!/bin/bash
# third pass: level heuristic (stop coarsening when the spheres are under-resolved), coarsest-level sweeps, sizes 256^3 .. 512^3 with and without multigrid
O=gpurun_out
timeout 150 python -m pytest tests/test_gpu_zz_poisson3d.py -x -q > $O/r2_mg_test.log 2>&1; tail -2 $O/r2_mg_test.log
run() {
    tag=$1; pre=$2; shift; shift
    env "$@" timeout 120 python tools/run_poisson3d.py --nx ${NX:-256} --precond $pre --repeat 2 > $O/r2_mg_$tag.json 2> $O/r2_mg_$tag.err
    python - "$tag" $O/r2_mg_$tag.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:28s} it {d['iterations']:4d}  loop {d['krylov_loop_ms']:8.1f} ms  {d['ms_per_iteration']:6.3f} ms/it  total {d['time_to_tolerance_ms']:8.1f} ms  launches {d['launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run 256_default mg A=1
run 256_sw24 mg PB200_MG_SWEEPS=24
run 256_sw6 mg PB200_MG_SWEEPS=6
run 256_res1 mg PB200_MG_RES=1.0
run 256_lev2_sw24 mg PB200_MG_LEVELS=2 PB200_MG_SWEEPS=24
NX=384 run 384_default mg A=1
NX=384 run 384_res1 mg PB200_MG_RES=1.0
NX=512 run 512_default mg A=1
NX=512 run 512_plain default A=1
