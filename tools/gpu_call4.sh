set -x
python -m pytest tests -q -m gpu 2>&1 | tail -5
for ce in 1 4 8 16; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --check-every $ce > gpurun_out/b_ce$ce.json 2>gpurun_out/b.err; python -c "
import json;d=json.load(open('gpurun_out/b_ce$ce.json'));print('ce',$ce,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['roofline']['avg_launch_us'])"; done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --method 1 > gpurun_out/b_cg.json 2>>gpurun_out/b.err; python -c "
import json;d=json.load(open('gpurun_out/b_cg.json'));print('cg',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])"
tail -5 gpurun_out/b.err
