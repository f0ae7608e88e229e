#!/bin/bash
# experiment: 3-D fused apply (1024 x 1024 x 128 diphasic): blocks per SM x pipeline depth, list order; DRAM bytes (ncu) and step times
O=gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
H="python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 1"
run() { tag=$1; shift; env "$@" PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 2 --log-file $O/r2_ncu_st_$tag.csv $H > $O/r2_ncu_st_$tag.log 2>&1; }
run index_1blk PB200_NO_BRICKS=1 PB200_DBG_F3=4
run index_S3 PB200_NO_BRICKS=1 PB200_F3_S=3
run index_S4 PB200_NO_BRICKS=1 PB200_F3_S=4
run index_S5 PB200_NO_BRICKS=1 PB200_F3_S=5
run runs_S4 PB200_F3_S=4
for v in "PB200_NO_BRICKS=1 PB200_F3_S=3" "PB200_NO_BRICKS=1 PB200_F3_S=4" "PB200_NO_BRICKS=1 PB200_F3_S=5" "PB200_NO_BRICKS=1 PB200_DBG_F3=4" "PB200_F3_S=4"; do
  echo "$v" >> $O/r2_h3d_stages.txt
  env $v python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 8 >> $O/r2_h3d_stages.txt 2>&1
done
