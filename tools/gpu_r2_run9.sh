#!/bin/bash
O=gpurun_out
python tests/checks/diag_virt.py > $O/r2_diag_virt.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu_c_host or virtual or subsample" > $O/r2_gputest9.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest9.log
cat $O/r2_diag_virt.txt; tail -30 $O/r2_gputest9.log
