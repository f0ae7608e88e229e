set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_pre.json 2> gpurun_out/bench_r1_pre.err; cat gpurun_out/bench_r1_pre.json; tail -3 gpurun_out/bench_r1_pre.err
timeout 600 python tools/run_heat3d.py --nx 256 --steps 20 > gpurun_out/heat3d_256.json 2> gpurun_out/heat3d.err; cat gpurun_out/heat3d_256.json; tail -3 gpurun_out/heat3d.err
timeout 900 python tools/run_heat3d.py --nx 512 --steps 20 > gpurun_out/heat3d_512.json 2>> gpurun_out/heat3d.err; cat gpurun_out/heat3d_512.json; tail -3 gpurun_out/heat3d.err
nvidia-smi --query-gpu=memory.used --format=csv
