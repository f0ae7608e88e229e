set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
for w in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --warm $w > gpurun_out/b13_w$w.json 2>gpurun_out/b13.err; python - <<PY
import json;d=json.load(open('gpurun_out/b13_w$w.json'));print('warm',$w,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
done
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --warm 2 > gpurun_out/b13_w2_100.json 2>gpurun_out/b13.err; python - <<PY
import json;d=json.load(open('gpurun_out/b13_w2_100.json'));print('warm 2, 100 steps',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
tail -3 gpurun_out/b13.err
