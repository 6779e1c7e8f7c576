#!/bin/bash
O=gpurun_out
bash tools/mac_sweep.sh > /dev/null 2>&1
python tests/checks/diag_c4_split.py > $O/r2_diag_c4.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_parity_fullsize.py::test_c4_automatic_partition_split_against_the_reference" > $O/r2_gputest2.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest2.log
: > $O/r2_graph_ab.txt
for CFG in "--shard-of 8" "--shard-of 4" "--shard-of 1" "--workload c2 --batch 1" "--workload c2" "--workload c4 --batch 1" "--workload c4"; do
  for NG in 0 1; do
    echo -n "$CFG nograph=$NG : " >> $O/r2_graph_ab.txt
    if [ $NG = 1 ]; then export BFCUDA_NO_GRAPH=1; else unset BFCUDA_NO_GRAPH; fi
    timeout 200 python bench.py --quick $CFG --steps 300 --warmup 5 2>> $O/r2_graph_ab.err | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f e2e %.0f step_us %.1f mac_us %.1f frac %.3f graph %d launches %d' % (q['value'], q['e2e_value'], q['ms_per_block'] * q['batch'] * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['uses_graph'], q['gpu_launches']))
except Exception as e:
    print('ERR', e)
" >> $O/r2_graph_ab.txt
  done
done
unset BFCUDA_NO_GRAPH
cat $O/r2_macsweep.txt $O/r2_diag_c4.txt $O/r2_graph_ab.txt; tail -5 $O/r2_gputest2.log
