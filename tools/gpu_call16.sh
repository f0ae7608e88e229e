set -x
CMD="python bench.py --steps 6 --warmup 5 --no-cpu --no-profile"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1e.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -n 1 gpurun_out/ncu1.log | cut -c1-200
