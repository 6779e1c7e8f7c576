#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_gputest3.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest3.log
bash tools/mac_sweep.sh > /dev/null 2>&1
: > $O/r2_graph_ab2.txt
for CFG in "--shard-of 8" "--shard-of 1" "--workload c2 --batch 1" "--workload c2" "--workload c4 --batch 1" "--workload c4"; do
  for G in 1 0; do
    echo -n "$CFG graph=$G : " >> $O/r2_graph_ab2.txt
    BFCUDA_GRAPH=$G timeout 200 python bench.py --quick $CFG --steps 300 --warmup 5 2>> $O/r2_graph_ab2.err | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f e2e %.0f step_us %.1f mac_us %.1f frac %.3f graph %d launches %d lat_ms %.3f' % (q['value'], q['e2e_value'], q['ms_per_block'] * q['batch'] * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['uses_graph'], q['gpu_launches'], q['sync_call_latency_ms']))
except Exception as e:
    print('ERR', e)
" >> $O/r2_graph_ab2.txt
  done
done
cat $O/r2_macsweep2.txt $O/r2_graph_ab2.txt; tail -30 $O/r2_gputest3.log
