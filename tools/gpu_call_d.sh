mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=25 -rf --durations=3 -k "rollout or specialised or golden_rollout" ) > gpurun_out/r2d_pytest.log 2>&1
tail -8 gpurun_out/r2d_pytest.log
timeout 300 python tools/ab_variants.py 65536 150 4096 > gpurun_out/r2d_ab.log 2>&1; cat gpurun_out/r2d_ab.log
timeout 300 python tools/warp_cycles.py 150 65536 > gpurun_out/r2d_warp_cycles.log 2>&1; cat gpurun_out/r2d_warp_cycles.log
