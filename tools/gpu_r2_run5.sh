#!/bin/bash
# experiments: batch groups for small shards; cp.async ring MAC for single blocks (c4); f64 fft2 tests + bench
O=gpurun_out
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
R=$O/r2_macsweep3.txt
: > $R
run() { # label, env..., args
  label=$1; shift
  echo -n "$label : " >> $R
  env "$@" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f step/mac %.2f split %d' % (q['value'], q['ms_per_block'] * q['batch'] * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['step_over_mac'], q['mac_split']))
except Exception as e:
    print('ERR', e)
" >> $R
}
for K in 8 4; do
  run "shard $K base W2 B8" BFCUDA_MAC_W=2 timeout 120 python bench.py --quick --shard-of $K --steps 200 --warmup 5
  run "shard $K groups2 W2 B4 tpb256" BFCUDA_MAC_W=2 BFCUDA_MAC_GROUPS=2 timeout 120 python bench.py --quick --shard-of $K --steps 200 --warmup 5
  run "shard $K groups2 W2 B4 tpb128" BFCUDA_MAC_W=2 BFCUDA_MAC_GROUPS=2 BFCUDA_MAC_TPB=128 timeout 120 python bench.py --quick --shard-of $K --steps 200 --warmup 5
  run "shard $K groups2 W4 B4" BFCUDA_MAC_W=4 BFCUDA_MAC_GROUPS=2 timeout 120 python bench.py --quick --shard-of $K --steps 200 --warmup 5
done
run "c4 B1 base" timeout 120 python bench.py --quick --workload c4 --batch 1 --steps 300 --warmup 5
for S in 8 12 16; do
  run "c4 B1 ring S$S" BFCUDA_MAC_B1_RING=1 BFCUDA_MAC_S=$S timeout 120 python bench.py --quick --workload c4 --batch 1 --steps 300 --warmup 5
done
run "c3 B1 base" timeout 120 python bench.py --quick --batch 1 --steps 100 --warmup 5
run "c3 B1 ring S8" BFCUDA_MAC_B1_RING=1 timeout 120 python bench.py --quick --batch 1 --steps 100 --warmup 5
unset BFCUDA_LIB BFCUDA_GRAPH
timeout 900 python -m pytest tests -m gpu -q -x -k "rs or diagonal or f64 or formats or mixing or link or bigfft or fullsize or batched" > $O/r2_gputest5.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest5.log
timeout 300 python bench.py --quick --workload c3f64 --steps 50 --warmup 3 > $O/r2_f64_b8.json 2> $O/r2_f64_b8.err
timeout 300 python bench.py --quick --workload c3f64 --batch 1 --steps 50 --warmup 3 > $O/r2_f64_b1.json 2> $O/r2_f64_b1.err
cat $R; tail -4 $O/r2_gputest5.log; cat $O/r2_f64_b8.json $O/r2_f64_b1.json | cut -c1-600
