#!/bin/bash
# Repeats the headline bench leg (no CPU baseline / DDIM / uint8 legs) and prints value, ms/step and the SM clock of each
# run: the B200 boxes run under sw_power_cap, single runs scatter by +-1 ms.   usage: tools/bench_repeat.sh [runs] [bench args]
runs=${1:-3}; shift
for i in $(seq 1 "$runs"); do
  python bench.py --no-cpu-baseline --no-ddim --no-uint8-leg "$@" 2>/dev/null > /tmp/bench_repeat.json
  python - <<'PY'
import json
d = json.load(open('/tmp/bench_repeat.json'))
print(round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 3), 'ms', 'e2e', round(d['e2e']['value'], 1), 'sm_mhz', d['clocks']['sm_mhz'])
PY
done
