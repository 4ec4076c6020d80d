# usage: bash tools/multi_gpu_call.sh N   (under gpurun --gpus N): sharded tests (N >= 2), weak + strong bench, phase trace
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rf --durations=3 ) > gpurun_out/r2m_pytest_sharded.log 2>&1; tail -5 gpurun_out/r2m_pytest_sharded.log
fi
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2m_weak_$N.json 2> gpurun_out/r2m_weak_$N.err; cut -c1-250 gpurun_out/r2m_weak_$N.json; tail -2 gpurun_out/r2m_weak_$N.err
timeout 600 $TR bench.py --gpus $N --steps 6 --warmup 3 --no-sweep --no-cpu-baseline --scaling strong --total-samples 1048576 > gpurun_out/r2m_strong_$N.json 2> gpurun_out/r2m_strong_$N.err; cut -c1-250 gpurun_out/r2m_strong_$N.json; tail -2 gpurun_out/r2m_strong_$N.err
MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py 65536 > gpurun_out/r2m_trace_weak_$N.log 2>&1; grep "trace rank 0" gpurun_out/r2m_trace_weak_$N.log | tail -1
MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py $((1048576 / N)) > gpurun_out/r2m_trace_strong_$N.log 2>&1; grep "trace rank 0" gpurun_out/r2m_trace_strong_$N.log | tail -1
