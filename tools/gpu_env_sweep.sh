#!/bin/bash
# Developer tool (run under gpurun): time the LM solve of the product build under launch-shape overrides.
cd "$(dirname "$0")/.."
for cfg in "" "EDSGPU_LEADER_CTAS=8" "EDSGPU_LEADER_CTAS=16" "EDSGPU_LEADER_CTAS=8 EDSGPU_EVAL_CTAS=128" "EDSGPU_LEADER_CTAS=8 EDSGPU_EVAL_CTAS=112" "EDSGPU_EVAL_CTAS=96"; do
  echo "== $cfg"
  for S in $1; do env $cfg timeout 40 python tools/lm_time.py $S 10 2>&1 | grep lm_time; done
done
