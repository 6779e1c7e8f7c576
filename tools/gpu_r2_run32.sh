#!/bin/bash
# step graphs block by block on every shard size
O=gpurun_out/r2_graph_b1.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 160 python bench.py --quick --shard-of $K --batch 1 --steps 1500 --warmup 30 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.0f step_us %.1f mac_us %.1f e2e %.0f lat %.3f graph %s' % (q['value'], q['ms_per_block'] * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('sync_call_latency_ms', 0), q.get('uses_graph')))
" >> $O
}
for K in 1 2 4 8; do
  for rep in 1 2; do
    q "rep $rep shard $K streams" BFCUDA_GRAPH=0
    q "rep $rep shard $K graph" BFCUDA_GRAPH=1
  done
done
cat $O
