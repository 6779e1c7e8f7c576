#!/bin/bash
# stream priorities on small shards: MAC first or FFT side first; with and without step graphs; repeats
O=gpurun_out/r2_prio.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 1000 --warmup 20 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f e2e %.0f graph %s' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('uses_graph')))
except Exception as e:
    print('ERR', e)
" >> $O
}
for K in 8 4 2; do
  for rep in 1 2; do
    q "rep $rep shard $K fft first, streams" BFCUDA_MAC_PRIO=0 BFCUDA_GRAPH=0
    q "rep $rep shard $K mac first, streams" BFCUDA_MAC_PRIO=1 BFCUDA_GRAPH=0
    q "rep $rep shard $K fft first, graph" BFCUDA_MAC_PRIO=0 BFCUDA_GRAPH=1
  done
done
cat $O
