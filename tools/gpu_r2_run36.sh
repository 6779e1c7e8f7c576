#!/bin/bash
# bisect BASELINE config 4 block by block over libraries of earlier commits, same harness
O=gpurun_out/r2_c4_after_revert.txt
: > $O
for LIB in tools/_build/wt_a265a37 .; do
  for I in 1; do
  echo -n "$LIB inline $I : " >> $O
  BFCUDA_MAC_INLINE_REDUCE=$I BFCUDA_LIB=$PWD/$LIB/brutefir_b200/libbfcuda.so timeout 100 python bench.py --workload c4 --quick --batch 1 --steps 3000 --warmup 100 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.1f step_us %.2f mac_us %.1f e2e %.0f graph %s launches %s' % (q['value'], q['ms_per_block'] * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('uses_graph'), q.get('gpu_launches')))
" >> $O
  done
done
cat $O
