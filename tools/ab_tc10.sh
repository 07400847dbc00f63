#!/bin/bash
set -u
O=gpurun_out/tc10; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "scorer or context or smoke" > $O/gputests.log 2>&1; rc=$?; echo "rc=$rc" >> $O/gputests.log
[ $rc -ne 0 ] && exit 0
python bench.py --workload attn_context --steps 150 --warmup 5 --no-cpu-baseline > $O/b_ctx_tc.json 2> $O/b_ctx_tc.err
MUAV_SCORER_TC=0 python bench.py --workload attn_context --steps 150 --warmup 5 --no-cpu-baseline > $O/b_ctx_fp32.json 2> $O/b_ctx_fp32.err
