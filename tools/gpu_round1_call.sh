mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -k "work_queue" ) > gpurun_out/pytest_queue.log 2>&1; echo "queue rc=$?" >> gpurun_out/pytest_queue.log
timeout 200 python tools/ab_variants.py > gpurun_out/ab.log 2>&1
timeout 100 python tools/warp_cycles.py 65536 12 > gpurun_out/warp_cycles_q12.log 2>&1
tail -3 gpurun_out/pytest_queue.log; cat gpurun_out/ab.log; tail -3 gpurun_out/warp_cycles_q12.log
