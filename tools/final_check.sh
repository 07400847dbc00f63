#!/bin/bash
# last GPU call of round 2: the whole GPU suite (with the CBBA bundle goldens), smoke(), the driver's short bench command
set -u
O=gpurun_out/final_r2; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 60 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/b_20.json 2> $O/b_20.err; echo "bench rc=$?" >> $O/b_20.err
tail -3 $O/gputests.log; tail -1 $O/smoke.log; cut -c1-300 $O/b_20.json
