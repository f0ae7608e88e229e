set -x
for i in 1 2 3; do timeout 300 python bench.py --no-cpu > gpurun_out/b20_$i.json 2>gpurun_out/b20.err; python - <<PY
import json;d=json.load(open('gpurun_out/b20_$i.json'));print('run',$i,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['clocks'])
PY
done
tail -3 gpurun_out/b20.err
