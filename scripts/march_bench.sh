#!/bin/bash
# marching kernels vs per-point kernels on BASELINE.json's configs[1..4] at full size (bench.py --workload)
mkdir -p gpurun_out
for wl in ${WORKLOADS:-cfg3 cfg2 cfg5 cfg4}; do
  steps=20; case $wl in cfg2|cfg5) steps=200;; esac
  for v in ${VARIANTS:-2 1}; do
    timeout 600 python bench.py --workload $wl --variant $v --steps $steps --warmup 3 --no-cpu > gpurun_out/bench_${wl}_v$v.json 2> gpurun_out/bench_${wl}_v$v.err
    echo "$wl v$v rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${wl}_v$v.json').read().strip().splitlines()[-1])
    r=d['roofline']; print('%.2f Gpt/s  %s  first %.3f ms second %.3f ms  dominant frac %.3f whole %.3f  e2e %.2f' % (d['value'], d['config']['kernels'], r['ms_first'], r['ms_second'], r['frac'], r['whole_step_frac'], d['e2e']['value']))
except Exception as e:
    print('FAILED', e)
PY
)"
  done
done
