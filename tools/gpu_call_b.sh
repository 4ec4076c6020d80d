mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracks_edges.py tests/test_gpu_loopback.py -m gpu -q --maxfail=25 -rf --durations=5 ) > gpurun_out/r2b_pytest.log 2>&1
tail -25 gpurun_out/r2b_pytest.log
timeout 300 python tools/ab_variants.py > gpurun_out/r2b_ab.log 2>&1; cat gpurun_out/r2b_ab.log
