for v in 1 0; do
  if [ $v = 1 ]; then export MUAV_NO_LEAN=1; else unset MUAV_NO_LEAN; fi
  python tools/kbench.py WPS_escort 8192 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nolean=$v escort kbench', round(d['single_step_flush_ms'],4), round(d['single_step_noflush_ms'],4), round(d['resident_150_ms_per_step'],4))"
  python tools/kbench.py WPS_commit 16384 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nolean=$v commit kbench', round(d['single_step_flush_ms'],4), round(d['single_step_noflush_ms'],4), round(d['resident_150_ms_per_step'],4))"
done
