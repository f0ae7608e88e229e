#!/bin/bash
# experiment: DRAM bytes of the 3-D fused apply (1024 x 1024 x 128 diphasic) under TMA L2-promotion settings and L2 eviction hints
O=gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
H="python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 1"
run() { tag=$1; shift; env "$@" PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 2 --log-file $O/r2_ncu_l2_$tag.csv $H > $O/r2_ncu_l2_$tag.log 2>&1; }
run promo0 PB200_TMA_PROMO=0
run promo3 PB200_TMA_PROMO=3
run hint1 PB200_L2HINT=1
run hint7 PB200_L2HINT=7
run hint6 PB200_L2HINT=6
run promo0_hint7 PB200_TMA_PROMO=0 PB200_L2HINT=7
run one_block_per_sm PB200_DBG_F3=4
