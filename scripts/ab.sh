#!/bin/bash
# A/B of kernel build variants / environment switches: bench at 1024^3 for each "lib[:ENV=VAL]" given
for spec in "$@"; do
  lib=${spec%%:*}; envs=""; [ "$spec" != "$lib" ] && envs=${spec#*:}
  echo "== $spec"
  env $envs WAVESIM_LIB=$PWD/wave-simulation_b200/csrc/$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu ${AB_ARGS} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['ms_velocity'], d['roofline']['ms_stress'], d['config']['kernels'])"
done
