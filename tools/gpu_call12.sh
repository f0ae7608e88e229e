set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-profile"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
PB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 600 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu1.log 2>&1
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf_apply_dense|kf_cg_update|kf_cg_p|kf_apply_band|kf_band_poly" -s 300 -c 10 -o gpurun_out/prof_cg_r1d $CMD > gpurun_out/ncu2.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; cat gpurun_out/bench_r1_d.json | cut -c1-400
