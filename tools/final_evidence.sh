#!/bin/bash
# Round-end evidence run on one B200 (called through gpurun): bench line, launch list, ncu full capture of the hot kernels, 3-D configs.
O=gpurun_out
python bench.py > $O/r1f_bench_n1.json 2> $O/r1f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r1f_bench_ref.json 2>> $O/r1f_bench_n1.err
python tools/run_heat3d.py --nx 512 --steps 50 > $O/r1f_heat3d_512.json 2>> $O/r1f_bench_n1.err
python tools/run_heat3d.py --nx 512 --steps 20 --diph > $O/r1f_heat3d_diph_512.json 2>> $O/r1f_bench_n1.err
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1f_launches.csv \
    python bench.py --no-cpu --no-profile --spinup 0 --warmup 8 --steps 3 > $O/r1f_ncu_launches.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf_apply_dense|kf_cg_update|kf_cg_p|kf_rhs_dense|kf_apply_band|kf_band_poly" \
    --launch-skip 420 -c 12 -o $O/r1f_hot_full python bench.py --no-cpu --no-profile --spinup 0 --warmup 8 --steps 3 > $O/r1f_ncu_full.log 2>&1
tail -2 $O/r1f_bench_n1.err
