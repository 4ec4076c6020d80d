mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:chol_cov -s 3 -c 1 -o gpurun_out/r2x_chol_cov -f python tools/profile_target.py 65536 1 > gpurun_out/r2x_ncu1.log 2>&1; tail -1 gpurun_out/r2x_ncu1.log
timeout 300 $NCU --set full --import-source on -k regex:ce_small_adapt -s 3 -c 1 -o gpurun_out/r2x_small_adapt -f python tools/profile_target.py 150 1 > gpurun_out/r2x_ncu2.log 2>&1; tail -1 gpurun_out/r2x_ncu2.log
