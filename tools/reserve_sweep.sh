#!/bin/bash
# Developer tool (run under gpurun): the main bench line for several numbers of SMs left to the event-frame builds
# (arguments: reserve[:ring_mb] ...; ring_mb = EDSGPU_ACC_RING_MB).
cd "$(dirname "$0")/.."
for a in "$@"; do
  r=${a%%:*}; m=${a#*:}; [ "$m" = "$a" ] && m=32
  EDSGPU_RESERVE_SMS=$r EDSGPU_ACC_RING_MB=$m python bench.py --steps 30 --warmup 3 --no-side --no-cpu-baseline --no-ba 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('reserve $r ring $m MB: %.1f windows/s, step %.4f ms, e2e %.1f, LM launch %.4f ms (frac %.3f), frames alone %.4f ms, shape %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'], d['event_frame']['ms'], d['config']['launch_shape']))"
done
