mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_loopback.py -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2p_pytest.log 2>&1
tail -5 gpurun_out/r2p_pytest.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 524288 2 2> gpurun_out/r2p_trace_512k.log; tail -1 gpurun_out/r2p_trace_512k.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 1048576 2 2> gpurun_out/r2p_trace_1m.log; tail -1 gpurun_out/r2p_trace_1m.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-sweep --no-cpu-baseline --scaling strong --total-samples 1048576 > gpurun_out/r2m_strong_1.json 2> gpurun_out/r2m_strong_1.err; cut -c1-300 gpurun_out/r2m_strong_1.json
