set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -q -m gpu 2>&1 | tail -40
python tools/debug_zero.py 2>&1 | tail -30
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo rc=$?; cat gpurun_out/bench_v1.json; tail -5 gpurun_out/bench_v1.err
python bench.py --steps 10 --warmup 3 --no-profile --no-cpu > gpurun_out/bench_v1_noprof.json 2>> gpurun_out/bench_v1.err; cat gpurun_out/bench_v1_noprof.json
