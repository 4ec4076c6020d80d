mkdir -p gpurun_out
for K in 262144 524288; do
  for cfg in "131072 8192" "131072 4096" "131072 2048" "1073741824 8192"; do
    set -- $cfg
    echo "K=$K big_thr=$1 keys=$2: $(MPOPIS_SELECT_BIG=$1 MPOPIS_SELECT_KEYS=$2 MPOPIS_TRACE=1 timeout 120 python tools/profile_target.py $K 2 2>&1 | grep -o 'select=[0-9.]*' | tail -1)"
  done
done 2>&1 | tee gpurun_out/r2ab_select.log
