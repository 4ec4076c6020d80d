mkdir -p gpurun_out
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2y_launches_C4.csv python tools/profile_c4.py C4 375 2 > gpurun_out/r2y_ncu1.log 2>&1; tail -2 gpurun_out/r2y_ncu1.log
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2y_launches_C3.csv python tools/profile_c4.py C3 4096 2 > gpurun_out/r2y_ncu2.log 2>&1; tail -2 gpurun_out/r2y_ncu2.log
