mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2l_pytest.log 2>&1
tail -8 gpurun_out/r2l_pytest.log
timeout 300 python tools/ab_variants.py 150 4096 8192 12288 > gpurun_out/r2l_ab.log 2>&1; cat gpurun_out/r2l_ab.log
timeout 300 python tools/warp_cycles.py 150 > gpurun_out/r2l_warp_cycles.log 2>&1; tail -6 gpurun_out/r2l_warp_cycles.log
