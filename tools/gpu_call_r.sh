mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_loopback.py -m gpu -q --maxfail=3 -rf --durations=5 ) > gpurun_out/r2r_pytest.log 2>&1
tail -30 gpurun_out/r2r_pytest.log | cut -c1-300
