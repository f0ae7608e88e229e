set -x
timeout 900 python tools/run_heat3d.py --nx 512 --steps 30 > gpurun_out/heat3d_512.json 2> gpurun_out/heat3d.err; cat gpurun_out/heat3d_512.json; tail -3 gpurun_out/heat3d.err
timeout 900 python tools/run_heat3d.py --nx 384 --steps 30 --diph > gpurun_out/heat3d_diph_384.json 2>> gpurun_out/heat3d.err; cat gpurun_out/heat3d_diph_384.json; tail -3 gpurun_out/heat3d.err
timeout 900 python tools/run_heat3d.py --nx 512 --steps 20 --diph > gpurun_out/heat3d_diph_512.json 2>> gpurun_out/heat3d.err; cat gpurun_out/heat3d_diph_512.json; tail -3 gpurun_out/heat3d.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
