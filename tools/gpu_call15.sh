set -x
python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; cat gpurun_out/bench_r1_e.json; tail -3 gpurun_out/bench_r1_e.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 | cut -c1-400
