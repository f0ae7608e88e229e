set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-profile"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kf_apply_dense|kf_cg_update|kf_cg_p|kf_apply_band" -s 40 -c 8 -o gpurun_out/prof_cg_r1b $CMD > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
