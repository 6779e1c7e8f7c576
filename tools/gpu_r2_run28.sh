#!/bin/bash
# the ordered sum of a split partition range inside k_mac (last block reduces): GPU suite, then BASELINE config 4 A/B
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_gputest.log 2>&1; echo rc=$? >> gpurun_out/r2d_gputest.log
O=gpurun_out/r2_inline_reduce_c4.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 160 python bench.py --workload $WL --quick --batch $B --steps 2000 --warmup 50 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.2f mac_us %.1f e2e %.0f split %s graph %s launches %s' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q.get('e2e_value', 0), q.get('mac_split'), q.get('uses_graph'), q.get('gpu_launches')))
except Exception as e:
    print('ERR', e)
" >> $O
}
for rep in 1 2; do
  WL=c4; B=1
  q "rep $rep c4 block by block, separate reduce kernel" BFCUDA_MAC_INLINE_REDUCE=0
  q "rep $rep c4 block by block, reduce inside k_mac" BFCUDA_MAC_INLINE_REDUCE=1
  WL=c2
  q "rep $rep c2 block by block, separate reduce kernel" BFCUDA_MAC_INLINE_REDUCE=0
  q "rep $rep c2 block by block, reduce inside k_mac" BFCUDA_MAC_INLINE_REDUCE=1
done
tail -3 gpurun_out/r2d_gputest.log; cat $O
