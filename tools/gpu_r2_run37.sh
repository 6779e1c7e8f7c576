#!/bin/bash
# partition split of the block-by-block MAC on the two short-partition BASELINE configurations
O=gpurun_out/r2_split_c2_c4.txt
: > $O
q() {
  echo -n "$WL split $1 : " >> $O
  BFCUDA_MAC_SPLIT=$1 timeout 100 python bench.py --workload $WL --quick --batch 1 --steps 3000 --warmup 100 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); q = d['quick']
print('value %.1f step_us %.2f mac_us %.1f e2e %.0f split %s launches %s' % (q['value'], q['ms_per_block'] * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('mac_split'), q.get('gpu_launches')))
" >> $O
}
WL=c2; for S in 1 2 4; do q $S; done
WL=c4; for S in 18 28 37 50 74; do q $S; done
cat $O
