mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=3 --deselect "tests/test_gpu_parity.py::test_baseline_configs_vs_oracle[car-3-cmamppi-375-50-10-kw3]" ) > gpurun_out/r2last_pytest.log 2>&1
tail -8 gpurun_out/r2last_pytest.log | cut -c1-250
( time timeout 900 python bench.py ) > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err; tail -3 gpurun_out/r2last_bench.err; cut -c1-400 gpurun_out/r2last_bench.json
