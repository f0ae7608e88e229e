set -x
CMD="python bench.py --steps 6 --warmup 5 --no-cpu --no-profile"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
PB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"kf_apply_dense|kf_cg_update|kf_cg_p|kf_apply_band|kf_band_poly" -s 200 -c 10 -o gpurun_out/prof_cg_r1f $CMD > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu2.log | cut -c1-200
