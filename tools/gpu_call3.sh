set -x
python -m pytest tests -q -m gpu 2>&1 | tail -30
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-profile"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --method 1 > gpurun_out/bench_r1_cg.json 2>> gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_cg.json
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 900 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kf_apply_dense -s 40 -c 3 -o gpurun_out/prof_apply_dense_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
