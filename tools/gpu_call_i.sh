mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_loopback.py tests/test_gpu_parity.py -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2i_pytest.log 2>&1
tail -12 gpurun_out/r2i_pytest.log
timeout 300 python tools/ab_variants.py 150 4096 8192 16384 > gpurun_out/r2i_ab.log 2>&1; cat gpurun_out/r2i_ab.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 3 2> gpurun_out/r2i_trace.log; tail -1 gpurun_out/r2i_trace.log
timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err; cut -c1-330 gpurun_out/r2i_bench.log; tail -3 gpurun_out/r2i_bench.err
