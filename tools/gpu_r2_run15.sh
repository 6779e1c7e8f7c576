#!/bin/bash
# what bounds the batched MAC on an 8-filter shard: memory side alone / arithmetic side alone (k_mac_tile MODE 1 / 2),
# and the SM clock during the kernel (ncu)
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
O=gpurun_out/r2_tilemodes.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch 8 --steps 200 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('step_us %.1f mac_us %.1f frac %.3f' % (q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac']))
except Exception as e:
    print('ERR', e)
" >> $O
}
for K in 8 1; do
  for G in 2 1; do
    for MODE in 0 1 2; do
      q "shard $K tile G $G mode $MODE" BFCUDA_MAC_TILE=1 BFCUDA_TILE_G=$G BFCUDA_TILE_TPG=64 BFCUDA_TILE_MODE=$MODE
    done
  done
done
M=sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum,dram__bytes_read.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__cycles_elapsed.avg.per_second
for T in 0 1; do
  BFCUDA_MAC_TILE=$T BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=64 ncu --metrics $M -k regex:k_mac -c 4 --clock-control none --csv --log-file gpurun_out/r2_clk_shard8_tile$T.csv python bench.py --quick --shard-of 8 --batch 8 --steps 3 --warmup 2 > /dev/null 2>&1
done
BFCUDA_MAC_TILE=0 ncu --metrics $M -k regex:k_mac -c 4 --clock-control none --csv --log-file gpurun_out/r2_clk_n1.csv python bench.py --quick --batch 8 --steps 3 --warmup 2 > /dev/null 2>&1
cat $O
