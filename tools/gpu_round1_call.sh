mkdir -p gpurun_out
( timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "tma_staged" ) > gpurun_out/pytest_tma.log 2>&1; echo "tma rc=$?" >> gpurun_out/pytest_tma.log
MPOPIS_APPLY_L=3 timeout 300 python -m pytest tests -m gpu -q --maxfail=12 -rf -k "golden_control_step or baseline_configs or every_policy or tiny_sizes or device_rng" > gpurun_out/pytest_apl3.log 2>&1
timeout 300 python tools/ab_variants.py > gpurun_out/ab.log 2>&1
MPOPIS_ROLLOUT_STAGE=1 timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -rf > gpurun_out/pytest_stage1.log 2>&1
MPOPIS_ROLLOUT_STAGE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_car_kernel -s 3 -c 1 -f -o gpurun_out/rollout_v4_tma python tools/profile_target.py 65536 1 > gpurun_out/ncu_full.log 2>&1
MPOPIS_APPLY_L=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:apply_L_dmma3 -s 2 -c 1 -f -o gpurun_out/apply_l3 python tools/profile_target.py 65536 1 > gpurun_out/ncu_full3.log 2>&1
tail -3 gpurun_out/pytest_tma.log; tail -2 gpurun_out/pytest_apl3.log; cat gpurun_out/ab.log; tail -2 gpurun_out/pytest_stage1.log
