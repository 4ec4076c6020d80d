# Final single-GPU call of round 2: launch lists, ncu --set full captures of the hot kernels, the default bench run,
# the reference arm, smoke(). Everything lands in gpurun_out/; summaries are written into profiles/ afterwards.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_K65536.csv python tools/profile_target.py 65536 2 > gpurun_out/r2_ncu_l1.log 2>&1; tail -1 gpurun_out/r2_ncu_l1.log
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_K150.csv python tools/profile_target.py 150 2 > gpurun_out/r2_ncu_l2.log 2>&1; tail -1 gpurun_out/r2_ncu_l2.log
timeout 400 $NCU --set full --import-source on -k regex:rollout_car_kernel -s 3 -c 1 -o gpurun_out/r2_rollout_v4_K65536 -f python tools/profile_target.py 65536 1 > gpurun_out/r2_ncu_f1.log 2>&1; tail -1 gpurun_out/r2_ncu_f1.log
timeout 400 $NCU --set full --import-source on -k regex:rollout_car_split -s 3 -c 1 -o gpurun_out/r2_rollout_split_K150 -f python tools/profile_target.py 150 1 > gpurun_out/r2_ncu_f2.log 2>&1; tail -1 gpurun_out/r2_ncu_f2.log
timeout 400 $NCU --set full --import-source on -k regex:rollout_car_split -s 3 -c 1 -o gpurun_out/r2_rollout_split_K16384 -f python tools/profile_target.py 16384 1 > gpurun_out/r2_ncu_f3.log 2>&1; tail -1 gpurun_out/r2_ncu_f3.log
timeout 400 $NCU --set full -k "regex:ce_select|chol_cov|apply_L_dmma2|syrk_dmma|elite_gather|ce_sums|shrink_q|scatter_reduce|philox_normals" -s 12 -c 12 -o gpurun_out/r2_side_K65536 -f python tools/profile_target.py 65536 1 > gpurun_out/r2_ncu_f4.log 2>&1; tail -1 gpurun_out/r2_ncu_f4.log
timeout 400 $NCU --set full -k "regex:ce_small_adapt" -s 1 -c 1 -o gpurun_out/r2_small_adapt_K150 -f python tools/profile_target.py 150 1 > gpurun_out/r2_ncu_f5.log 2>&1; tail -1 gpurun_out/r2_ncu_f5.log
ls -la gpurun_out/r2_*.ncu-rep
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -3 gpurun_out/r2_bench_final.err; cut -c1-600 gpurun_out/r2_bench_final.json
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -3 gpurun_out/r2_bench_reference.err; cut -c1-400 gpurun_out/r2_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
