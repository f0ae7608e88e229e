set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python tools/run_heat3d.py --nx 384 --steps 30 --diph > gpurun_out/heat3d_diph_384.json 2> gpurun_out/heat3d.err; cat gpurun_out/heat3d_diph_384.json; tail -3 gpurun_out/heat3d.err
