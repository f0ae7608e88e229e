set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --no-cpu > gpurun_out/b24.json 2>gpurun_out/b24.err; python - <<PY
import json;d=json.load(open('gpurun_out/b24.json'));print('bench',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],[ (k['kernel'][:12],round(k['avg_launch_us'],1),round(k['frac'],2)) for k in d['roofline']['kernels']])
PY
tail -3 gpurun_out/b24.err
timeout 900 python tools/run_heat3d.py --nx 512 --steps 30 > gpurun_out/heat3d_512.json 2> gpurun_out/heat3d.err; cat gpurun_out/heat3d_512.json; tail -3 gpurun_out/heat3d.err
