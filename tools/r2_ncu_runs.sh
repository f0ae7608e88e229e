#!/bin/bash
# experiment: DRAM bytes / time of the 3-D fused apply (1024 x 1024 x 128 diphasic) for run shapes (y tiles, z tiles) of the block-wise tile assignment
O=gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
H="python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 1"
run() { tag=$1; shift; env "$@" PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 2 --log-file $O/r2_ncu_run_$tag.csv $H > $O/r2_ncu_run_$tag.log 2>&1; }
run 2x8 PB200_RUN=2,8
run 1x8 PB200_RUN=1,8
run 1x4 PB200_RUN=1,4
run 2x4 PB200_RUN=2,4
run 4x8 PB200_RUN=4,8
run 2x16 PB200_RUN=2,16
python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 10 > $O/r2_h3d_run_2x8.json 2>&1
PB200_RUN=1,8 python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 10 > $O/r2_h3d_run_1x8.json 2>&1
PB200_RUN=4,8 python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 10 > $O/r2_h3d_run_4x8.json 2>&1
PB200_NO_BRICKS=1 python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 10 > $O/r2_h3d_run_off.json 2>&1
