mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf --durations=5 ) > gpurun_out/r2z_pytest.log 2>&1
tail -14 gpurun_out/r2z_pytest.log | cut -c1-250
timeout 200 python tools/profile_c4.py C4 375 4 2>&1 | tail -3
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2z_launches_C4.csv python tools/profile_c4.py C4 375 2 > gpurun_out/r2z_ncu1.log 2>&1; tail -1 gpurun_out/r2z_ncu1.log
