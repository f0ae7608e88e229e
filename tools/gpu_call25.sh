set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in 4 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --no-cpu > gpurun_out/b25_n$n.json 2> gpurun_out/b25_n$n.err; echo rc=$?; python - <<PY
import json;d=json.load(open('gpurun_out/b25_n$n.json'));print('N',$n,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['config']['grid'])
PY
done
timeout 200 python bench.py --no-cpu > gpurun_out/b25_n1.json 2> gpurun_out/b25_n1.err; python - <<PY
import json;d=json.load(open('gpurun_out/b25_n1.json'));print('N',1,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
