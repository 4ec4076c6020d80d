# round 2, call A: the new selection / loop-back paths first, then the whole GPU suite, smoke, a short bench + launch list
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_loopback.py -m gpu -q --maxfail=40 -rf -x --durations=5 ) > gpurun_out/r2a_pytest_new.log 2>&1
tail -15 gpurun_out/r2a_pytest_new.log
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=5 --deselect tests/test_gpu_select.py --deselect tests/test_gpu_loopback.py ) > gpurun_out/r2a_pytest.log 2>&1
tail -8 gpurun_out/r2a_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.log; tail -3 gpurun_out/r2a_bench.err
MPOPIS_CE_SELECT=0 timeout 300 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/r2a_bench_sort.log 2>> gpurun_out/r2a_bench.err; cut -c1-300 gpurun_out/r2a_bench_sort.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2a_launches.csv python tools/profile_target.py 65536 3 > gpurun_out/r2a_ncu.log 2>&1; tail -2 gpurun_out/r2a_ncu.log
