#!/bin/bash
# developer sweep: ring depth of the marching kernels on the BASELINE configs
mkdir -p gpurun_out
for wl in ${WORKLOADS:-cfg3 cfg4 cfg2 cfg5}; do
  steps=20; case $wl in cfg2|cfg5) steps=200;; esac
  for nst in ${STAGES:-0 2 3 4}; do
    WS_MARCH_STAGES=$nst timeout 600 python bench.py --workload $wl --variant 2 --steps $steps --warmup 3 --no-cpu > gpurun_out/sw_${wl}_$nst.json 2> gpurun_out/sw_${wl}_$nst.err
    echo "$wl nst=$nst chunk=${WS_MARCH_CHUNK:-auto} rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sw_${wl}_$nst.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('%.2f Gpt/s first %.3f ms second %.3f ms  dominant frac %.3f whole %.3f' % (d['value'], r['ms_first'], r['ms_second'], r['frac'], r['whole_step_frac']))
except Exception as e:
    print('FAILED', e)
PY
)"
  done
done
