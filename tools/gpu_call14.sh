set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
for w in 3 4 5; do timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --warm $w > gpurun_out/b14_w$w.json 2>gpurun_out/b14.err; python - <<PY
import json;d=json.load(open('gpurun_out/b14_w$w.json'));print('warm',$w,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['config']['final_rel_residual'])
PY
done
tail -3 gpurun_out/b14.err
