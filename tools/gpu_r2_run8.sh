#!/bin/bash
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2_gputest8.log 2>&1; echo "pytest rc=$?" >> $O/r2_gputest8.log
tail -40 $O/r2_gputest8.log
