#!/bin/bash
# new defaults (two bins per thread + whole partition sums on small shards, cooperative kernel for 16-block launches):
# the GPU suite, the shard series, the bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gputest.log 2>&1; echo rc=$? >> gpurun_out/r2b_gputest.log
O=gpurun_out/r2b_shards.txt
: > $O
for B in 8 16; do
for K in 8 4 2 1; do
  echo -n "B $B shard $K default library : " >> $O
  timeout 120 python bench.py --quick --shard-of $K --batch $B --steps 300 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f frac %.3f split %s graph %s' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q.get('mac_split'), q.get('uses_graph')))
except Exception as e:
    print('ERR', e)
" >> $O
done
done
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_gputest.log; cat $O
