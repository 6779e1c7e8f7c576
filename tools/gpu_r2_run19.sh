#!/bin/bash
# k_mac_coop: memory side alone (MODE 1) and arithmetic side alone (MODE 2) on an 8-filter shard and the full job
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
O=gpurun_out/r2_coopmodes.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('mac_us %.1f' % (q['mac_kernel_ms'] * 1e3))
except Exception as e:
    print('ERR', e)
" >> $O
}
for K in 8 1; do
  for V in "2 32" "1 64" "4 32"; do
    set -- $V
    for MODE in 0 1 2; do
      q "shard $K coop G $1 TPG $2 mode $MODE" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2 BFCUDA_TILE_MODE=$MODE
    done
  done
done
cat $O
