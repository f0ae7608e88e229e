set -x
python -m pytest tests -q -m gpu 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --no-cpu > gpurun_out/bf_$i.json 2>gpurun_out/bf.err; python - <<PY
import json;d=json.load(open('gpurun_out/bf_$i.json'));print('run',$i,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['config']['rhs_assembly_ms_per_step'],d['config']['solve_ms_per_step'])
PY
done
tail -3 gpurun_out/bf.err
