#!/bin/bash
# A/B with repeats: batched MAC on an 8-filter shard (8 blocks per call), and 16 blocks per call on shards of 8 / 4 / 2
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
O=gpurun_out/r2_coop_ab.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch $B --steps 1000 --warmup 20 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac']))
except Exception as e:
    print('ERR', e)
" >> $O
}
B=8; K=8
for rep in 1 2 3; do
  q "rep $rep B 8 shard 8 batch2" BFCUDA_MAC_TILE=0
  q "rep $rep B 8 shard 8 coop G 2 TPG 64" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=64
  q "rep $rep B 8 shard 8 coop G 1 TPG 64" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=1 BFCUDA_TILE_TPG=64
done
B=16
for K in 8 4 2; do
  for rep in 1 2; do
    q "rep $rep B 16 shard $K batch2" BFCUDA_MAC_TILE=0
    q "rep $rep B 16 shard $K coop G 2 TPG 64" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=64
    q "rep $rep B 16 shard $K coop G 2 TPG 32" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=32
  done
done
cat $O
