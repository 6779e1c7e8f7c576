#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_gputest12.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest12.log
: > $O/r2_graph_copies.txt
for CFG in "--workload c2 --batch 1" "--workload c2" "--workload c4 --batch 1" "--workload c4"; do
  for GC in 1 0; do
    echo -n "$CFG copies_in_graph=$GC : " >> $O/r2_graph_copies.txt
    BFCUDA_GRAPH_COPIES=$GC timeout 200 python bench.py --quick $CFG --steps 300 --warmup 5 2>> $O/r2_graph_copies.err | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f e2e %.0f step_us %.1f graph %d lat_ms %.3f' % (q['value'], q['e2e_value'], q['ms_per_block'] * q['batch'] * 1e3, q['uses_graph'], q['sync_call_latency_ms']))
except Exception as e:
    print('ERR', e)
" >> $O/r2_graph_copies.txt
  done
done
BFCUDA_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_mac|k_split_reduce" -s 40 -c 2 -f -o $O/r2_c4_mac python bench.py --quick --workload c4 --batch 1 --steps 30 --warmup 3 > $O/r2_c4_ncu.log 2>&1
cat $O/r2_graph_copies.txt; tail -25 $O/r2_gputest12.log
