#!/bin/bash
# A/B of one environment switch inside running frames on ONE box (boxes of the pool differ by 7-9 %):
#   bash tools/ab_env.sh TDNET_TC_PAIR_MODE plain tail [reps]
# alternates the values, prints frames/s, ms per frame, end-to-end frames/s and the two in-frame kernel probes per run.
cd "$(dirname "$0")/.."
VAR=$1; A=$2; B=$3; REPS=${4:-2}
for rep in $(seq $REPS); do
  for v in "$A" "$B"; do
    echo -n "$VAR=$v  "
    env "$VAR=$v" timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-fast-mode --sustain-seconds 0.5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'frames/s', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), ' conv', round(d['roofline']['ms_per_launch'],4), 'attn', round(d['roofline_attention']['ms_per_launch'],4))"
  done
done
