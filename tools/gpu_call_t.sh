mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=8 ) > gpurun_out/r2t_pytest.log 2>&1
tail -25 gpurun_out/r2t_pytest.log | cut -c1-250
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 150 3 2> gpurun_out/r2t_trace_150.log; tail -1 gpurun_out/r2t_trace_150.log
timeout 300 python tools/ab_small.py > gpurun_out/r2t_ab.log 2>&1; tail -6 gpurun_out/r2t_ab.log | cut -c1-200
