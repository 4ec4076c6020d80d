mkdir -p gpurun_out
timeout 300 python tools/ablate.py 150 > gpurun_out/r2g_ablate.log 2>&1; cat gpurun_out/r2g_ablate.log
