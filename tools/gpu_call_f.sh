mkdir -p gpurun_out
timeout 300 python tools/warp_cycles.py 150 65536 > gpurun_out/r2f_warp_cycles.log 2>&1; cat gpurun_out/r2f_warp_cycles.log
