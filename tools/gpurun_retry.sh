# usage: bash tools/gpurun_retry.sh <log> <gpus> <timeout-s> '<command>'   — retries while the pod answers busy/transient
LOG=$1; GPUS=$2; TMO=$3; CMD=$4
for i in $(seq 1 20); do
  if [ "$GPUS" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $TMO -- "$CMD" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TMO -- "$CMD" > $LOG 2>&1; fi
  if grep -q "status=transient\|status=busy\|nothing was charged" $LOG; then sleep 200; continue; fi
  break
done
tail -40 $LOG | cut -c1-300
