mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tracks_edges.py tests/test_gpu_loopback.py tests/test_gpu_select.py -m gpu -q --maxfail=25 -rf --durations=5 ) > gpurun_out/r2c_pytest.log 2>&1
tail -12 gpurun_out/r2c_pytest.log
timeout 300 python tools/ab_variants.py 65536 150 4096 > gpurun_out/r2c_ab.log 2>&1; cat gpurun_out/r2c_ab.log
timeout 300 python tools/warp_cycles.py 150 65536 > gpurun_out/r2c_warp_cycles.log 2>&1; cat gpurun_out/r2c_warp_cycles.log
MPOPIS_GRAPH_VERBOSE=1 timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench.err; cut -c1-400 gpurun_out/r2c_bench.log; tail -3 gpurun_out/r2c_bench.err
MPOPIS_GRAPH=0 timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2c_bench_nograph.log 2>> gpurun_out/r2c_bench.err; cut -c1-400 gpurun_out/r2c_bench_nograph.log
