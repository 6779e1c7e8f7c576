#!/bin/bash
# batched MAC block size at 16 / 32 / 64 filters under the final defaults (MAC stream first)
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so
O=gpurun_out/r2_macsweep_tpb_final.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 400 --warmup 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.0f step_us %.1f mac_us %.1f frac %.3f e2e %.0f' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q.get('e2e_value', 0)))
" >> $O
}
for K in 4 2 1; do
  for TPB in 256 128 1128 192; do
    q "shard $K W 2 S 8 TPB $TPB" BFCUDA_MAC_W=2 BFCUDA_MAC_S=8 BFCUDA_MAC_TPB=$TPB BFCUDA_MAC_SPLIT=1
  done
done
cat $O
