#!/bin/bash
# A/B of one environment switch inside ONE GPU session (same box): tools/ab_env.sh VAR valueA valueB [pairs]
var=$1; a=$2; b=$3; pairs=${4:-2}
for i in $(seq 1 "$pairs"); do
  for v in "$a" "$b"; do
    env "$var=$v" python bench.py --no-cpu-baseline --no-ddim --no-uint8-leg 2>/dev/null > /tmp/ab_env.json
    python - "$var=$v" <<'PY'
import json, sys
d = json.load(open('/tmp/ab_env.json'))
print(sys.argv[1], round(d['value'], 1), 'samples/s', round(d['ms_per_step'], 3), 'ms', 'sm_mhz', d['clocks']['sm_mhz'])
PY
  done
done
