#!/bin/bash
# two GPUs: the NCCL split-output path inside bench.py, and the two-GPU parity test
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo2.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus" > $O/r2_gputest7.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest7.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err; echo "bench rc=$?" >> $O/r2_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 2 > $O/r2_ref_n2.json 2> $O/r2_ref_n2.err
tail -5 $O/r2_gputest7.log; grep -i "nranks\|error\|Traceback" $O/r2_bench_n2.err | head -20; tail -3 $O/r2_bench_n2.err; cut -c1-400 $O/r2_bench_n2.json
