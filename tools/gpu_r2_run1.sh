#!/bin/bash
# round-2 GPU session 1: parity suite, new bench line, small-shard MAC A/B, light FFT A/B, host topology
mkdir -p gpurun_out
O=gpurun_out
(nvidia-smi topo -m; lscpu | head -30; cat /sys/devices/system/node/node*/cpulist; nvidia-smi --query-gpu=index,pci.bus_id --format=csv) > $O/r2_topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_gputest1.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest1.log
timeout 600 python bench.py --steps 100 --warmup 5 > $O/r2_bench1.json 2> $O/r2_bench1.err; echo "bench rc=$?" >> $O/r2_bench1.err
for K in 8 4 2; do
  timeout 300 python bench.py --shard-of $K --steps 200 --warmup 5 --no-extras --no-cpu-baseline > $O/r2_shard${K}_new.json 2> $O/r2_shard${K}_new.err
  BFCUDA_MAC_NARROW_MAX_BINS=0 timeout 300 python bench.py --shard-of $K --steps 200 --warmup 5 --no-extras --no-cpu-baseline > $O/r2_shard${K}_old.json 2> $O/r2_shard${K}_old.err
done
BFCUDA_FFT2_LIGHT=1 timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > $O/r2_fftlight.json 2> $O/r2_fftlight.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > $O/r2_fftbase.json 2> $O/r2_fftbase.err
tail -3 $O/r2_gputest1.log
