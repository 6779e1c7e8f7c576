#!/bin/bash
# stream priorities at the full job (64 filters on one GPU) and block by block
O=gpurun_out/r2_prio_n1.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 160 python bench.py --quick --shard-of $K --batch $B --steps 600 --warmup 20 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f e2e %.0f lat %.3f graph %s' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('sync_call_latency_ms', 0), q.get('uses_graph')))
except Exception as e:
    print('ERR', e)
" >> $O
}
K=1
for B in 8 1; do
  for rep in 1 2; do
    q "rep $rep B $B full job fft first, streams" BFCUDA_MAC_PRIO=0 BFCUDA_GRAPH=0
    q "rep $rep B $B full job mac first, streams" BFCUDA_MAC_PRIO=1 BFCUDA_GRAPH=0
    q "rep $rep B $B full job fft first, graph" BFCUDA_MAC_PRIO=0 BFCUDA_GRAPH=1
    q "rep $rep B $B full job mac first, graph" BFCUDA_MAC_PRIO=1 BFCUDA_GRAPH=1
  done
done
K=8; B=1
q "B 1 shard 8 fft first, streams" BFCUDA_MAC_PRIO=0 BFCUDA_GRAPH=0
q "B 1 shard 8 mac first, streams" BFCUDA_MAC_PRIO=1 BFCUDA_GRAPH=0
q "B 1 shard 8 auto" BFCUDA_MAC_PRIO=0
cat $O
