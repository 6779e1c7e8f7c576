#!/bin/bash
# experiment: k_mac_tile (block-cooperative bulk-copy staging, block groups) against k_mac_batch2 on shards of
# 8 / 4 / 2 / 1 ranks' worth of filters.  BFCUDA_MAC_TILE=1 selects the tile kernel, BFCUDA_TILE_G / _TPG the variant.
export BFCUDA_GRAPH=0
O=${1:-gpurun_out/r2_tilesweep.txt}
BATCHES=${BATCHES:-8}
: > $O
run() {   # label, env...
  echo -n "$1 : " >> $O
  shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch $B --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f step/mac %.2f split %s' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['step_over_mac'], q.get('mac_split')))
except Exception as e:
    print('ERR', e)
" >> $O
}
for B in $BATCHES; do
for K in ${SHARDS:-8 4 2 1}; do
  run "B $B shard $K batch2 (baseline)" BFCUDA_MAC_TILE=0
  for V in ${VARIANTS:-"1 64" "1 128" "2 32" "2 64" "2 128" "4 32" "4 64"}; do
    set -- $V
    run "B $B shard $K tile G $1 TPG $2" BFCUDA_MAC_TILE=1 BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2
  done
done
done
cat $O
