#!/bin/bash
# k_pack with the single-precision quantiser fast path: GPU suite, kernel time, step time
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gputest.log 2>&1; echo rc=$? >> gpurun_out/r2e_gputest.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_pack\|k_unpack -c 8 --csv --log-file gpurun_out/r2e_pack.csv python bench.py --quick --steps 3 --warmup 2 > /dev/null 2>&1
O=gpurun_out/r2e_quick.txt
: > $O
for K in 1 8; do
  for rep in 1 2; do
  echo -n "shard $K rep $rep : " >> $O
  timeout 160 python bench.py --quick --shard-of $K --batch 8 --steps 600 --warmup 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.0f step_us %.1f mac_us %.1f e2e %.0f step/mac %.2f' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q['step_over_mac']), d.get('stage_ms_per_block'))
" >> $O
  done
done
tail -3 gpurun_out/r2e_gputest.log; cat $O; grep -E "k_pack|k_unpack" gpurun_out/r2e_pack.csv | cut -d, -f5,13- | head -12
