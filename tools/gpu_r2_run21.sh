#!/bin/bash
# ncu --set full with source counters: the batched MAC of an 8-filter shard, three kernels
export BFCUDA_LIB=$PWD/brutefir_b200/libbfcuda_sweep.so BFCUDA_GRAPH=0
cap() {   # name, env...
  n=$1; shift
  env "$@" ncu --set full --import-source on --clock-control none -k regex:k_mac -c 1 -s 4 -f -o gpurun_out/r2s_$n python bench.py --quick --shard-of 8 --batch 8 --steps 4 --warmup 2 > gpurun_out/r2s_$n.log 2>&1
}
cap batch2_w1 BFCUDA_MAC_TILE=0
cap batch2_w2 BFCUDA_MAC_TILE=0 BFCUDA_MAC_W=2 BFCUDA_MAC_S=8 BFCUDA_MAC_TPB=256 BFCUDA_MAC_SPLIT=1
cap coop_g1 BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=1 BFCUDA_TILE_TPG=64
cap coop_g2 BFCUDA_MAC_TILE=2 BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=64
ls -la gpurun_out/r2s_*
