mkdir -p gpurun_out
./tools/bin/chain_bench > gpurun_out/r2e_chain_bench.txt 2>&1; cat gpurun_out/r2e_chain_bench.txt
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loopback.py -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2e_pytest.log 2>&1
tail -6 gpurun_out/r2e_pytest.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 4 2> gpurun_out/r2e_trace.log; tail -2 gpurun_out/r2e_trace.log
timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err; cut -c1-330 gpurun_out/r2e_bench.log; tail -3 gpurun_out/r2e_bench.err
