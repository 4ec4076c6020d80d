mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracks_edges.py tests/test_external_env.py -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2h_pytest.log 2>&1
tail -12 gpurun_out/r2h_pytest.log
timeout 300 python tools/ab_variants.py 150 4096 16384 28416 65536 > gpurun_out/r2h_ab.log 2>&1; cat gpurun_out/r2h_ab.log
timeout 300 python tools/warp_cycles.py 150 4096 > gpurun_out/r2h_warp_cycles.log 2>&1; cat gpurun_out/r2h_warp_cycles.log
