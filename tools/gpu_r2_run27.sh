#!/bin/bash
# batched MAC: two partition steps per wait (PAIR = 2)
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so
O=gpurun_out/r2_macsweep_pair.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f split %s' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q.get('mac_split')))
except Exception as e:
    print('ERR', e)
" >> $O
}
for K in 8 4 2 1; do
  for V in "8 256" "8 1256" "8 3256" "8 4256" "12 3256"; do
    set -- $V
    q "shard $K W 2 S $1 TPB $2" BFCUDA_MAC_TILE=0 BFCUDA_MAC_W=2 BFCUDA_MAC_S=$1 BFCUDA_MAC_TPB=$2 BFCUDA_MAC_SPLIT=1
  done
done
cat $O
