#!/bin/bash
# batched MAC, two bins per thread: thread-block sizes that deal the threads evenly over 148 SMs
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
O=gpurun_out/r2_macsweep_tpb.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f step/mac %.2f' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['step_over_mac']))
except Exception as e:
    print('ERR', e)
" >> $O
}
for K in 8 4 2 1; do
  q "shard $K default" BFCUDA_MAC_TILE=0
  for TPB in 256 224 192 160 128 96 448; do
    q "shard $K W 2 S 8 TPB $TPB" BFCUDA_MAC_W=2 BFCUDA_MAC_S=8 BFCUDA_MAC_TPB=$TPB
  done
done
cat $O
