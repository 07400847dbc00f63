#!/bin/bash
# the reference's own regression scenarios (tests/test_reference_regressions.py) on the CUDA facade
set -u
O=gpurun_out/final_r2; mkdir -p $O
timeout 58 python -m pytest tests/test_reference_regressions.py -m gpu -x -q > $O/gputests_regressions.log 2>&1; echo "rc=$?" >> $O/gputests_regressions.log
tail -5 $O/gputests_regressions.log
