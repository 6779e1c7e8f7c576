#!/bin/bash
# final evidence: ncu launch list of the bench command, full captures of the default MAC on an 8-filter shard and at N = 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_launches_bench_final.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_mac_batch2 -c 1 -s 4 -f -o gpurun_out/r2f_shard8_mac python bench.py --quick --shard-of 8 --batch 8 --steps 4 --warmup 2 > gpurun_out/r2f_shard8_mac.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_mac_coop -c 1 -s 4 -f -o gpurun_out/r2f_shard8_coop16 python bench.py --quick --shard-of 8 --batch 16 --steps 4 --warmup 2 > gpurun_out/r2f_shard8_coop16.log 2>&1
ncu --set full --clock-control none -k regex:k_mac -c 1 -s 8 -f -o gpurun_out/r2f_c4_mac python bench.py --workload c4 --quick --batch 1 --steps 6 --warmup 3 > gpurun_out/r2f_c4_mac.log 2>&1
ls -la gpurun_out/r2f_* gpurun_out/r2_launches_bench_final.csv
