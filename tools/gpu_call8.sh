set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b8_n2.json 2> gpurun_out/b8_n2.err; tail -3 gpurun_out/b8_n2.err; cat gpurun_out/b8_n2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/b8_n1.json 2>gpurun_out/b8_n1.err; cat gpurun_out/b8_n1.json
