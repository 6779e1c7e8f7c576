#!/bin/bash
O=gpurun_out
R=$O/r2_run13.txt
: > $R
q() { label=$1; shift; echo -n "$label : " >> $R; env "$@" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); q = d['quick']; s = d['stage_ms_per_block']
    print('value %.0f e2e %.0f step_us %.1f mac_us %.1f frac %.3f graph %d lat_ms %.3f | fwd %.1f inv %.1f us/blk' % (q['value'], q['e2e_value'], q['ms_per_block'] * q['batch'] * 1e3, q['mac_kernel_ms'] * 1e3, q['roofline_frac'], q['uses_graph'], q['sync_call_latency_ms'], s['forward']*1e3, s['inverse']*1e3))
except Exception as e:
    print('ERR', e)
" >> $R; }
q "c4 B1" timeout 200 python bench.py --quick --workload c4 --batch 1 --steps 300 --warmup 5
q "c4 B8" timeout 200 python bench.py --quick --workload c4 --steps 300 --warmup 5
q "c3 rw8" timeout 200 python bench.py --quick --steps 200 --warmup 5
q "c3 rw16" BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_rw16.so timeout 200 python bench.py --quick --steps 200 --warmup 5
q "c3 rw8 again" timeout 200 python bench.py --quick --steps 200 --warmup 5
q "c3 rw16 again" BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_rw16.so timeout 200 python bench.py --quick --steps 200 --warmup 5
q "shard8 rw8" timeout 200 python bench.py --quick --shard-of 8 --steps 200 --warmup 5
q "shard8 rw16" BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_rw16.so timeout 200 python bench.py --quick --shard-of 8 --steps 200 --warmup 5
timeout 900 python -m pytest tests -m gpu -q -x -k "engine or fullsize or powersave" > $O/r2_gputest13.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest13.log
cat $R; tail -4 $O/r2_gputest13.log
