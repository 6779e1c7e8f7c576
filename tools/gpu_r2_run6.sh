#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_gputest6.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest6.log
timeout 900 python bench.py --steps 100 --warmup 5 > $O/r2_bench6.json 2> $O/r2_bench6.err; echo "bench rc=$?" >> $O/r2_bench6.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke6.log 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_launches_ncu.log 2>&1
# the bulk-copy (TMA-staged) single-block MAC next to the direct-load one
BFCUDA_MAC_VARIANT=1 BFCUDA_GRAPH=0 python bench.py --quick --batch 1 --steps 50 --warmup 3 > $O/r2_tma_plain.json 2>&1
BFCUDA_MAC_VARIANT=1 BFCUDA_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_mac" -s 20 -c 1 -f -o $O/r2_mac_tma python bench.py --quick --batch 1 --steps 20 --warmup 3 > $O/r2_tma_ncu.log 2>&1
BFCUDA_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"k_mac" -s 20 -c 1 -f -o $O/r2_mac_b1 python bench.py --quick --batch 1 --steps 20 --warmup 3 > $O/r2_b1_ncu.log 2>&1
tail -15 $O/r2_gputest6.log; cat $O/r2_bench6.err | tail -5; cat $O/r2_smoke6.log | tail -2; cat $O/r2_tma_plain.json | cut -c1-300
