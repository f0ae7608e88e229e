set -x
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/b11_n2.json 2> gpurun_out/b11_n2.err; tail -2 gpurun_out/b11_n2.err; python - <<PY
import json;d=json.load(open('gpurun_out/b11_n2.json'));print('N2 plain',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
PB200_GRAPH_NCCL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/b11_n2g.json 2> gpurun_out/b11_n2g.err; tail -2 gpurun_out/b11_n2g.err; python - <<PY
import json;d=json.load(open('gpurun_out/b11_n2g.json'));print('N2 graph',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
