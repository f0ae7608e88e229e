set -x
python -m pytest tests -q -m gpu 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --no-cpu > gpurun_out/b20_$i.json 2>gpurun_out/b20.err; python - <<PY
import json;d=json.load(open('gpurun_out/b20_$i.json'));print('run',$i,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],[ (k['kernel'][:12],round(k['avg_launch_us'],1)) for k in d['roofline']['kernels']])
PY
done
tail -3 gpurun_out/b20.err
