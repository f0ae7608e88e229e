O=gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
H="python tools/run_heat3d.py --diph --nx 1024 --nz 128 --steps 1"
PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 4 --log-file $O/r2_ncu_apply_1024x128_bricks.csv $H > $O/r2_ncu_a1.log 2>&1
PB200_NO_BRICKS=1 PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 4 --log-file $O/r2_ncu_apply_1024x128_index.csv $H > $O/r2_ncu_a2.log 2>&1
PB200_BRICK=4,8,4 PB200_NO_GRAPH=1 ncu $M -k regex:kf3_apply --launch-skip 12 -c 4 --log-file $O/r2_ncu_apply_1024x128_b484.csv $H > $O/r2_ncu_a3.log 2>&1
grep -h kf3_apply $O/r2_ncu_apply_1024x128_*.csv | cut -c1-20,150-400 | head -40
