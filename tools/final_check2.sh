#!/bin/bash
# market-allocator and facade GPU tests on the last build (bundle-4 CBBA golden, backend hook)
set -u
O=gpurun_out/final_r2; mkdir -p $O
timeout 110 python -m pytest tests -m gpu -x -q -k "cbba or pi or facade or replay" > $O/gputests_market.log 2>&1; echo "rc=$?" >> $O/gputests_market.log
tail -3 $O/gputests_market.log
