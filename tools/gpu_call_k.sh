mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ce_select -s 4 -c 2 -o gpurun_out/r2k_select_cluster -f python tools/profile_target.py 65536 2 > gpurun_out/r2k_ncu1.log 2>&1; tail -2 gpurun_out/r2k_ncu1.log
MPOPIS_SELECT_CLUSTER=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ce_select -s 4 -c 2 -o gpurun_out/r2k_select_coop -f python tools/profile_target.py 65536 2 > gpurun_out/r2k_ncu2.log 2>&1; tail -2 gpurun_out/r2k_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_car_split -s 2 -c 1 -o gpurun_out/r2k_split_k4096 -f python tools/profile_target.py 4096 1 > gpurun_out/r2k_ncu3.log 2>&1; tail -2 gpurun_out/r2k_ncu3.log
ls -la gpurun_out/*.ncu-rep | tail -5
