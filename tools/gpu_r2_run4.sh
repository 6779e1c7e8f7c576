#!/bin/bash
# ncu: the batched MAC of an 8-filter shard (what one rank of an 8-GPU run executes), full set + source
O=gpurun_out
python bench.py --quick --shard-of 8 --steps 20 --warmup 3 > $O/r2_s8_plain.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_mac_batch2 -s 30 -c 2 -f -o $O/r2_shard8_mac python bench.py --quick --shard-of 8 --steps 20 --warmup 3 > $O/r2_s8_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mac_batch2 -s 30 -c 1 -f -o $O/r2_n1_mac python bench.py --quick --steps 20 --warmup 3 > $O/r2_n1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_forward2|k_inverse2|k_pack|k_unpack" -s 40 -c 4 -f -o $O/r2_n1_fft python bench.py --quick --steps 20 --warmup 3 > $O/r2_n1_fft_ncu.log 2>&1
python -m pytest tests/test_gpu_engine.py -q -k "errors_and_limits" > $O/r2_gputest4.log 2>&1
tail -3 $O/r2_s8_ncu.log $O/r2_gputest4.log; ls -la $O/*.ncu-rep
