mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_loopback.py -m gpu -q --maxfail=6 -rf --durations=5 -k "rng or early_stop" ) > gpurun_out/r2r_pytest0.log 2>&1
tail -6 gpurun_out/r2r_pytest0.log | cut -c1-300
( time timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q --maxfail=6 -rf --durations=5 ) > gpurun_out/r2r_pytest.log 2>&1
tail -12 gpurun_out/r2r_pytest.log | cut -c1-300
