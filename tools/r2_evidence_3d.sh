#!/bin/bash
# ncu full capture of the 3-D diphasic iteration kernels (320^3, one B200); the band-polynomial kernel is in r2_ncu_full_3d_bandpoly_raw.csv
O=gpurun_out
H="python tools/run_heat3d.py --diph --nx 320 --steps 2"
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf3_apply|kf2_update|kf_apply_band" --launch-skip 8 -c 9 -o $O/r2_full_3d_iter $H > $O/r2_ncu_full_3d_iter.log 2>&1
ncu -i $O/r2_full_3d_iter.ncu-rep --page raw --csv > $O/r2_ncu_full_3d_iter_raw.csv 2>/dev/null
# in-library checks instead of compute-sanitizer (closed on this pool): bounds checks of every tile access, staged box vs global memory, header vs record
PB200_DBG_F3=16 python -m pytest tests/test_gpu_fastpath_parity.py -m gpu -x -q -k "fused_pipelined or fused_band_launches" > $O/r2_dbg_bounds_checks.log 2>&1
tail -n 3 $O/r2_dbg_bounds_checks.log
