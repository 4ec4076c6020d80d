# usage: bash tools/multi_gpu_quick.sh N tag — peer-memory transport only: sharded tests (N = 2), weak bench, phase trace
N=${1:-2}; TAG=${2:-r2u}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -rf -k peer ) > gpurun_out/${TAG}_pytest_sharded.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_sharded.log
fi
timeout 600 $TR bench.py --gpus $N --no-sweep --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_weak_peer_$N.json 2> gpurun_out/${TAG}_weak_peer_$N.err
grep '^{' gpurun_out/${TAG}_weak_peer_$N.json | cut -c1-330; tail -1 gpurun_out/${TAG}_weak_peer_$N.err | cut -c1-200
MPOPIS_TRACE=1 timeout 300 $TR tools/trace_sharded.py 65536 > gpurun_out/${TAG}_trace_weak_peer_$N.log 2>&1; grep "trace rank 0" gpurun_out/${TAG}_trace_weak_peer_$N.log | tail -1
if [ "$N" != "2" ]; then
  timeout 600 $TR bench.py --gpus $N --no-sweep --no-cpu-baseline --steps 6 --warmup 3 --scaling strong --total-samples 1048576 > gpurun_out/${TAG}_strong_peer_$N.json 2> gpurun_out/${TAG}_strong_peer_$N.err
  grep '^{' gpurun_out/${TAG}_strong_peer_$N.json | cut -c1-330
fi
