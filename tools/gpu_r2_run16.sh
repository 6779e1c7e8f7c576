#!/bin/bash
# k_mac_coop: correctness through the MAC-bearing tests, then the sweep on shards of 8 / 4 / 2 / 1 ranks
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
L=gpurun_out/r2_coop_tests.log
: > $L
for V in "2 32 16" "4 32 16" "2 64 16" "1 32 16"; do
  set -- $V
  echo "== coop G $1 TPG $2 S $3" >> $L
  BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2 BFCUDA_TILE_S=$3 timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4 >> $L
done
O=gpurun_out/r2_coopsweep.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch $B --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f step/mac %.2f' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['step_over_mac']))
except Exception as e:
    print('ERR', e)
" >> $O
}
B=8
for K in 8 4 2 1; do
  q "shard $K batch2 (baseline)" BFCUDA_MAC_TILE=0
  for V in "2 32 16" "2 32 32" "2 64 16" "2 64 32" "4 32 16" "4 32 32" "4 64 16" "1 32 16" "1 64 16"; do
    set -- $V
    q "shard $K coop G $1 TPG $2 S $3" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2 BFCUDA_TILE_S=$3
  done
done
K=8
q "shard 8 coop G 2 TPG 32 mode 1 (memory side)" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=32 BFCUDA_TILE_MODE=1
q "shard 8 coop G 2 TPG 32 mode 2 (arithmetic side)" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=32 BFCUDA_TILE_MODE=2
K=1
q "shard 1 coop G 2 TPG 32 mode 1 (memory side)" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=32 BFCUDA_TILE_MODE=1
q "shard 1 coop G 2 TPG 32 mode 2 (arithmetic side)" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=32 BFCUDA_TILE_MODE=2
B=16
for K in 8 1; do
  q "B 16 shard $K batch2 (baseline)" BFCUDA_MAC_TILE=0
  for V in "2 32" "4 32" "2 64" "4 64"; do
    set -- $V
    q "B 16 shard $K coop G $1 TPG $2" BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2
  done
done
cat $L $O
