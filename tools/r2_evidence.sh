#!/bin/bash
# Round-2 evidence run on one B200 (called through gpurun): launch lists and ncu full captures of the hot kernels -- the 2-D bench config
# (2048^2 diphasic BE) and the one-GPU slab of the north-star config (1024 x 1024 x 128 diphasic BE) -- and the in-library bounds checks
# (compute-sanitizer is closed on this pool).  Bench numbers are never taken under ncu.
O=gpurun_out
B="python bench.py --no-cpu --no-profile --no-3d --spinup 0 --warmup 8 --steps 3"
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_2d.csv $B > $O/r2_ncu_launches_2d.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf3_apply|kf2_update_b" --launch-skip 120 -c 4 -o $O/r2_full_2d $B > $O/r2_ncu_full_2d.log 2>&1
ncu -i $O/r2_full_2d.ncu-rep --page raw --csv > $O/r2_ncu_full_2d_raw.csv 2>/dev/null; rm -f $O/r2_full_2d.ncu-rep   # (gpurun brings back at most 64 MiB)
H="python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 1"
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 380 -c 400 --csv --log-file $O/r2_launches_3d_1024x128.csv $H > $O/r2_ncu_launches_3d.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf3_apply|kf2_update|kf_apply_band" --launch-skip 9 -c 6 -o $O/r2_full_3d $H > $O/r2_ncu_full_3d.log 2>&1
ncu -i $O/r2_full_3d.ncu-rep --page raw --csv > $O/r2_ncu_full_3d_raw.csv 2>/dev/null; rm -f $O/r2_full_3d.ncu-rep
PB200_NO_GRAPH=1 ncu --set full --clock-control none -k regex:kf_band_poly --launch-skip 120 -c 1 -o $O/r2_full_3d_bandpoly $H > $O/r2_ncu_full_3d_bandpoly.log 2>&1
ncu -i $O/r2_full_3d_bandpoly.ncu-rep --page raw --csv > $O/r2_ncu_full_3d_bandpoly_raw.csv 2>/dev/null; rm -f $O/r2_full_3d_bandpoly.ncu-rep
cuobjdump -sass penguin.jl_b200/libpenguin_b200.so 2>/dev/null | grep -E "Function : .*kf3_apply|UTMALDG|SYNCS" | awk '/Function/{f=$0} /UTMALDG/{u[f]++} /SYNCS/{s[f]++} END{for(k in u) print k, "UTMALDG", u[k], "SYNCS", s[k]}' > $O/r2_sass_tma_grep.txt
PB200_DBG_F3=16 python -m pytest tests/test_gpu_fastpath_parity.py -m gpu -x -q -k "fused_pipelined or fused_band_launches" > $O/r2_dbg_bounds_checks.log 2>&1
tail -n 3 $O/r2_dbg_bounds_checks.log
