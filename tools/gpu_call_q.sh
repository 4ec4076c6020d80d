mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_loopback.py tests/test_gpu_parity.py -m gpu -q --maxfail=25 -rf --durations=3 -k "not cmamppi-375" ) > gpurun_out/r2q_pytest.log 2>&1
tail -5 gpurun_out/r2q_pytest.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 3 2> gpurun_out/r2q_trace_64k.log; tail -1 gpurun_out/r2q_trace_64k.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 1048576 2 2> gpurun_out/r2q_trace_1m.log; tail -1 gpurun_out/r2q_trace_1m.log
