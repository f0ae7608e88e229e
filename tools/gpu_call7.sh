set -x
python -m pytest tests -q -m gpu 2>&1 | tail -5
PB200_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b7.json 2>gpurun_out/b7.err; grep "band block" gpurun_out/b7.err | head -2; python - <<PY
import json;d=json.load(open('gpurun_out/b7.json'));print('prec',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['roofline']['avg_launch_us'])
PY
PB200_NO_BAND_PREC=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b7n.json 2>>gpurun_out/b7.err; python - <<PY
import json;d=json.load(open('gpurun_out/b7n.json'));print('noprec',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
grep -v pb200 gpurun_out/b7.err | tail -5
