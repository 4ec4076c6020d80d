mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_parity.py -m gpu -q --maxfail=25 -rf --durations=3 ) > gpurun_out/r2j_pytest.log 2>&1
tail -8 gpurun_out/r2j_pytest.log
timeout 300 python tools/ab_variants.py 150 4096 8192 12288 > gpurun_out/r2j_ab.log 2>&1; cat gpurun_out/r2j_ab.log
timeout 300 python tools/warp_cycles.py 150 > gpurun_out/r2j_warp_cycles.log 2>&1; cat gpurun_out/r2j_warp_cycles.log
MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 3 2> gpurun_out/r2j_trace.log; tail -1 gpurun_out/r2j_trace.log
MPOPIS_SELECT_CLUSTER=0 MPOPIS_TRACE=1 timeout 200 python tools/profile_target.py 65536 3 2> gpurun_out/r2j_trace_coop.log; tail -1 gpurun_out/r2j_trace_coop.log
