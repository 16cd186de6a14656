#!/bin/bash
# A/B harness used throughout round 1: graph-replay bench of the current build against debug knobs
# (TAVSR_DEBUG="k=v,...", TAVSR_* environment switches) on ONE box, because only whole-replay
# timings decide (DESIGN.md "What did NOT pay").  usage: gpurun -- 'bash tools/ab.sh "11=1" "12=1"'
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
python bench.py --no-cpu 2>&1 | tail -1 | p default
for knob in "$@"; do
  TAVSR_DEBUG="$knob" python bench.py --no-cpu 2>&1 | tail -1 | p "debug[$knob]"
done
