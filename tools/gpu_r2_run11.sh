#!/bin/bash
# eight GPUs: the scaling line with the copy-only ceiling and the NCCL split-output check
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo8.txt 2>&1; nproc >> $O/r2_topo8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err; echo "bench rc=$?" >> $O/r2_bench_n8.err
tail -3 $O/r2_bench_n8.err; cut -c1-300 $O/r2_bench_n8.json
