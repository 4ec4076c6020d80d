mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=5 ) > gpurun_out/pytest.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
