mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=25 -rf --durations=3 -k "rollout or specialised or golden or baseline_configs or every_policy" ) > gpurun_out/r2n_pytest.log 2>&1
tail -6 gpurun_out/r2n_pytest.log
timeout 300 python tools/ab_variants.py 150 4096 8192 12288 > gpurun_out/r2n_ab.log 2>&1; cat gpurun_out/r2n_ab.log
timeout 300 python tools/warp_cycles.py 150 > gpurun_out/r2n_warp_cycles.log 2>&1; tail -6 gpurun_out/r2n_warp_cycles.log
