#!/bin/bash
# correctness of k_mac_tile through the GPU suite's MAC-bearing tests, then the sweep
export BFCUDA_MAC_TILE=1
for V in "2 64" "4 32" "1 64"; do
  set -- $V
  echo "== tile G $1 TPG $2" >> gpurun_out/r2_tile_tests.log
  BFCUDA_TILE_G=$1 BFCUDA_TILE_TPG=$2 timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 >> gpurun_out/r2_tile_tests.log
done
BFCUDA_TILE_G=2 BFCUDA_TILE_TPG=64 timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -x -q -k c3 2>&1 | tail -5 >> gpurun_out/r2_tile_tests.log
unset BFCUDA_MAC_TILE
tools/tile_sweep.sh gpurun_out/r2_tilesweep.txt
BATCHES=16 SHARDS="8 1" VARIANTS="2:64 4:32 4:64" true
