for rep in 1 2; do
for v in 1 0; do
  if [ $v = 1 ]; then export MUAV_NO_LEAN=1; else unset MUAV_NO_LEAN; fi
  python tools/kbench.py WPS_hard 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nolean=$v kbench', round(d['single_step_flush_ms'],4), round(d['single_step_noflush_ms'],4), round(d['resident_150_ms_per_step'],4))"
  python bench.py --steps 450 --warmup 150 --cpu-seconds 0 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nolean=$v bench', round(d['value']/1e6,3), round(d['roofline']['kernel_ms_per_launch'],4), round(d['ms_per_step'],4), d['error_flags'])"
done; done
