# usage: bash tools/multi_gpu_call.sh N [tag]  (under gpurun --gpus N): sharded tests (N = 2), weak + strong bench with the
# peer-memory collectives and with NCCL (MPOPIS_COMM_PEER=0), phase traces
N=${1:-2}; TAG=${2:-r2s}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  ( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rf --durations=3 ) > gpurun_out/${TAG}_pytest_sharded.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_sharded.log
fi
run() {  # name, extra env, bench args
  env $2 timeout 600 $TR bench.py --gpus $N --no-sweep --no-cpu-baseline $3 > gpurun_out/${TAG}_$1_$N.json 2> gpurun_out/${TAG}_$1_$N.err
  grep '^{' gpurun_out/${TAG}_$1_$N.json | cut -c1-330; tail -1 gpurun_out/${TAG}_$1_$N.err | cut -c1-200
}
run weak_peer "MPOPIS_COMM_PEER=1" "--steps 10 --warmup 3"
run weak_nccl "MPOPIS_COMM_PEER=0" "--steps 10 --warmup 3"
run strong_peer "MPOPIS_COMM_PEER=1" "--steps 6 --warmup 3 --scaling strong --total-samples 1048576"
if [ "$N" != "4" ]; then run strong_nccl "MPOPIS_COMM_PEER=0" "--steps 6 --warmup 3 --scaling strong --total-samples 1048576"; fi
MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py 65536 > gpurun_out/${TAG}_trace_weak_peer_$N.log 2>&1; grep "trace rank 0" gpurun_out/${TAG}_trace_weak_peer_$N.log | tail -1
MPOPIS_COMM_PEER=0 MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py 65536 > gpurun_out/${TAG}_trace_weak_nccl_$N.log 2>&1; grep "trace rank 0" gpurun_out/${TAG}_trace_weak_nccl_$N.log | tail -1
MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py $((1048576 / N)) > gpurun_out/${TAG}_trace_strong_peer_$N.log 2>&1; grep "trace rank 0" gpurun_out/${TAG}_trace_strong_peer_$N.log | tail -1
