mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=8 ) > gpurun_out/pytest.log 2>&1
timeout 200 python tools/ab_variants.py > gpurun_out/ab.log 2>&1
MPOPIS_APPLY_L=2 timeout 300 python -m pytest tests -m gpu -q --maxfail=12 -rf -k "golden_control_step or baseline_configs or every_policy or tiny_sizes or device_rng" > gpurun_out/pytest_apl2.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py 65536 3 > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_car_kernel -s 3 -c 1 -f -o gpurun_out/rollout_v4c python tools/profile_target.py 65536 1 > gpurun_out/ncu_full.log 2>&1
MPOPIS_APPLY_L=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:apply_L_dmma2 -s 2 -c 1 -f -o gpurun_out/apply_l2c python tools/profile_target.py 65536 1 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/ab.log; tail -3 gpurun_out/pytest_apl2.log; cat gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
