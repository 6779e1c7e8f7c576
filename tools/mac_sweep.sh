#!/bin/bash
# experiment: ring depth / threads per block / lanes of the batched MAC at B = 8, on shards of 1, 2, 4, 8 ranks
# (needs brutefir_b200/libbfcuda_sweep.so: bf_mac_batch.cu built with -DBF_MAC_SWEEP)
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
O=gpurun_out/r2_macsweep2.txt
: > $O
for K in 8 4 1; do
  for V in "2 8 256" "2 10 256" "2 12 256" "2 12 128" "2 16 128" "2 16 256" "2 24 256" "1 8 64" "1 16 64" "1 24 64"; do
    set -- $V
    echo -n "shard $K W $1 S $2 TPB $3 : " >> $O
    BFCUDA_MAC_W=$1 BFCUDA_MAC_S=$2 BFCUDA_MAC_TPB=$3 timeout 120 python bench.py --quick --shard-of $K --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f step/mac %.2f' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['step_over_mac']))
except Exception as e:
    print('ERR', e)
" >> $O
  done
done
unset BFCUDA_LIB BFCUDA_GRAPH
