mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=5 ) > gpurun_out/pytest.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k150.csv python tools/profile_target.py 150 2 > gpurun_out/ncu_launch150.log 2>&1
tail -4 gpurun_out/pytest.log; cat gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err
