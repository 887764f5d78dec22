#!/bin/bash
# Same-box A/B of the headline: tools/gpu_ab.sh "<label>:<ENV=1 ...>" ...   (label "base" with no env = default build)
Q="--steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-decode"
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  [ "$envs" = "$spec" ] && envs=""
  out=$(env $envs python bench.py $Q 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])')
  echo "AB $label [$envs]: $out"
done
