mkdir -p gpurun_out
for i in 1 2; do timeout 150 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x 2>&1 | tail -2; done | tee gpurun_out/r2flake.log
