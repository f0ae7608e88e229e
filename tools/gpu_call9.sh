set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-profile"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 700 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -n 2 gpurun_out/ncu1.log | cut -c1-300
