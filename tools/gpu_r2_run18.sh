#!/bin/bash
# slim FFT-side kernels (64 registers, 2 blocks per SM, tables through L1): do they overlap the batched MAC better?
export BFCUDA_GRAPH=0
O=gpurun_out/r2_fftslim.txt
: > $O
q() {
  echo -n "$1 : " >> $O; shift
  env "$@" timeout 120 python bench.py --quick --shard-of $K --batch $B --steps 300 --warmup 10 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']
    print('value %.0f step_us %.1f mac_us %.1f step/mac %.2f stages %s' % (q['value'], q['ms_per_block'] * $B * 1e3, q['mac_kernel_ms'] * 1e3, q['step_over_mac'], json.dumps(d.get('stage_ms_per_block'))))
except Exception as e:
    print('ERR', e)
" >> $O
}
for B in 8 16; do
for K in 1 2 8; do
  q "B $B shard $K fft default" BFCUDA_FFT_SLIM=0
  q "B $B shard $K fft slim" BFCUDA_FFT_SLIM=1
done
done
timeout 300 env BFCUDA_FFT_SLIM=1 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity_fullsize.py -m gpu -x -q -k "c3" 2>&1 | tail -3 >> $O
cat $O
