#!/bin/bash
# A/B of environment knobs (AB_CONFIGS="base A=1 A=1:B=2"): prints value + writes stage tables per configuration.
# AB_ARGS: extra bench.py arguments (default: the fp32 tensor-core mode only, no oracle / reference legs)
set -u
mkdir -p gpurun_out
ARGS=${AB_ARGS:---precision fp32 --no-secondary --no-parity --no-cpu-baseline}
for cfg in ${AB_CONFIGS:-base}; do
  tag=$(echo $cfg | tr '=:' '__')
  if [ "$cfg" = "base" ]; then env_cmd=""; else env_cmd="env $(echo $cfg | tr ':' ' ')"; fi
  $env_cmd timeout 300 python bench.py --steps 10 --warmup 3 $ARGS --stage-table gpurun_out/ab_${tag}.csv > gpurun_out/ab_${tag}.log 2>&1
  echo "== $cfg: $(tail -1 gpurun_out/ab_${tag}.log | cut -c1-110)"
done
