#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_gputest10.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest10.log
timeout 900 python bench.py --steps 200 --warmup 5 > $O/r2_bench10.json 2> $O/r2_bench10.err; echo "bench rc=$?" >> $O/r2_bench10.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $O/r2_ref10.json 2> $O/r2_ref10.err
tail -12 $O/r2_gputest10.log; tail -3 $O/r2_bench10.err
