mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=5 ) > gpurun_out/r2w_pytest.log 2>&1
tail -14 gpurun_out/r2w_pytest.log | cut -c1-250
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 3 2> gpurun_out/r2w_trace_64k.log; tail -1 gpurun_out/r2w_trace_64k.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 150 3 2> gpurun_out/r2w_trace_150.log; tail -1 gpurun_out/r2w_trace_150.log
timeout 300 python tools/ab_small.py 150 > gpurun_out/r2w_ab.log 2>&1; tail -3 gpurun_out/r2w_ab.log
timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; cut -c1-330 gpurun_out/r2w_bench.json
