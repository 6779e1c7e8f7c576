#!/bin/bash
# the full job, 8 blocks per call: one MAC block per SM (shared-memory pad) + slim FFT blocks beside it
O=gpurun_out/r2_corun.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 160 python bench.py --quick --batch $B --steps 400 --warmup 10 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f e2e %.0f step/mac %.2f' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q['step_over_mac']))
except Exception as e:
    print('ERR', e)
" >> $O
}
B=8
for rep in 1 2; do
  q "rep $rep default" BFCUDA_GRAPH=0
  q "rep $rep slim fft only" BFCUDA_GRAPH=0 BFCUDA_FFT_SLIM=1
  q "rep $rep mac 1 block/SM (pad 52 KB)" BFCUDA_GRAPH=0 BFCUDA_MAC_SMEM_PAD=53248
  q "rep $rep mac 1 block/SM + slim fft" BFCUDA_GRAPH=0 BFCUDA_MAC_SMEM_PAD=53248 BFCUDA_FFT_SLIM=1
  q "rep $rep mac 1 block/SM + slim fft, fft first" BFCUDA_GRAPH=0 BFCUDA_MAC_SMEM_PAD=53248 BFCUDA_FFT_SLIM=1 BFCUDA_MAC_PRIO=0
  q "rep $rep mac 1 block/SM + slim fft, graph" BFCUDA_GRAPH=1 BFCUDA_MAC_SMEM_PAD=53248 BFCUDA_FFT_SLIM=1
done
cat $O
