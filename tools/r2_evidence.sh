#!/bin/bash
# Round-2 evidence run on one B200 (called through gpurun): launch lists, ncu full captures of the hot kernels (2-D bench config and a 3-D
# diphasic case), compute-sanitizer memcheck / racecheck of the fused path.  Bench numbers are never taken under ncu / the sanitizer.
O=gpurun_out
B="python bench.py --no-cpu --no-profile --no-3d --spinup 0 --warmup 8 --steps 3"
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_2d.csv $B > $O/r2_ncu_launches_2d.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf3_apply|kf2_update_b" --launch-skip 120 -c 4 -o $O/r2_full_2d $B > $O/r2_ncu_full_2d.log 2>&1
ncu -i $O/r2_full_2d.ncu-rep --page raw --csv > $O/r2_ncu_full_2d_raw.csv 2>/dev/null
H="python tools/run_heat3d.py --diph --nx 320 --steps 2"
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2_launches_3d_diph320.csv $H > $O/r2_ncu_launches_3d.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf3_apply|kf2_update|kf_apply_band|kf_band_poly" --launch-skip 100 -c 6 -o $O/r2_full_3d $H > $O/r2_ncu_full_3d.log 2>&1
ncu -i $O/r2_full_3d.ncu-rep --page raw --csv > $O/r2_ncu_full_3d_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fastpath_parity.py -m gpu -x -q -k "test_diph_256_interior_tiles_vs_oracle and BE and 8 and (fused_pipelined or fused_band_launches)" > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck exit $?" >> $O/r2_sanitizer_racecheck_smoke.log
tail -3 $O/r2_sanitizer_memcheck.log $O/r2_sanitizer_racecheck_smoke.log
