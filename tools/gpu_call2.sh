set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -40
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo rc=$?; cat gpurun_out/bench_v2.json; tail -5 gpurun_out/bench_v2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-profile --no-cpu --check-every 4 > gpurun_out/bench_v2b.json 2>> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2b.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-profile --no-cpu --method 1 > gpurun_out/bench_v2cg.json 2>> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2cg.json
