mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_loopback.py -m gpu -q --maxfail=25 -rf --durations=3 -k "not cmamppi-375" ) > gpurun_out/r2o_pytest.log 2>&1
tail -6 gpurun_out/r2o_pytest.log
timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2o_bench.log 2> gpurun_out/r2o_bench.err; cut -c1-330 gpurun_out/r2o_bench.log; tail -3 gpurun_out/r2o_bench.err
timeout 300 python tools/ab_variants.py 16384 18944 24576 > gpurun_out/r2o_ab.log 2>&1; cat gpurun_out/r2o_ab.log
