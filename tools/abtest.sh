python -m pytest tests -m gpu -x -q -k "lean_and_general" 2>&1 | tail -3
run() { python tools/kbench.py WPS_hard 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['single_step_flush_ms'],4), round(d['single_step_noflush_ms'],4), round(d['resident_150_ms_per_step'],4))"; }
for m in 31 0 1 3 7 15 23 27 29 30; do MUAV_SYNC_MASK=$m run "mask=$m"; done
for w in 2 3 4 5; do MUAV_CTA_WARPS=$w run "warps=$w"; done
for w in 3 4; do MUAV_CTA_WARPS=$w MUAV_SYNC_MASK=0 run "warps=$w mask=0"; done
