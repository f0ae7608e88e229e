set -x
python -m pytest tests -q -m gpu 2>&1 | tail -5
for m in 0 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --method $m > gpurun_out/b5_m$m.json 2>gpurun_out/b5.err; python - <<PY
import json;d=json.load(open('gpurun_out/b5_m$m.json'));print('m',$m,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['roofline'])
PY
done
tail -5 gpurun_out/b5.err
