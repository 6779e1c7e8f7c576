#!/bin/bash
# packed S24_LE tile path: tests, then the headline shape with massive_config's own sample format
python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "packed_s24 or sample_formats or formats" > gpurun_out/r2g_tests.log 2>&1; echo rc=$? >> gpurun_out/r2g_tests.log
O=gpurun_out/r2_s24le.txt
: > $O
for WL in c3 c3s24le; do
  echo -n "$WL : " >> $O
  timeout 200 python bench.py --workload $WL --quick --batch 8 --steps 400 --warmup 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.0f step_us %.1f mac_us %.1f e2e %.0f step/mac %.2f' % (q['value'], q['ms_per_block'] * 8e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q['step_over_mac']), d.get('stage_ms_per_block'))
" >> $O
done
tail -4 gpurun_out/r2g_tests.log; cat $O
