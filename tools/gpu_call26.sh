set -x
for n in 4; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --no-cpu > gpurun_out/b25_n$n.json 2> gpurun_out/b25_n$n.err; echo rc=$?; python - <<PY
import json;d=json.load(open('gpurun_out/b25_n$n.json'));print('N',$n,d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'],d['config']['grid'])
PY
done
