set -x
PB200_DEBUG=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_parity.py > gpurun_out/mgp.log 2>&1; echo rc=$?; grep "state\|OK\|Error\|error\|peer-memory\|FAIL" gpurun_out/mgp.log | head -20
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu > gpurun_out/p2p_n2.json 2> gpurun_out/p2p_n2.err; echo rc=$?; python - <<PY
import json;d=json.load(open('gpurun_out/p2p_n2.json'));print('N2 p2p',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
PB200_NO_P2P=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --no-cpu > gpurun_out/nccl_n2.json 2> gpurun_out/nccl_n2.err; echo rc=$?; python - <<PY
import json;d=json.load(open('gpurun_out/nccl_n2.json'));print('N2 nccl',d['value'],d['ms_per_step'],d['config']['iters_per_step'],d['e2e']['value'])
PY
tail -3 gpurun_out/p2p_n2.err | cut -c1-200
