#!/bin/bash
# 2-D tile kernels: parity tests, then cfg2 / cfg5 with both tile heights, the staging-only ceiling, and the marching kernels beside them
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "tile2d or default_kernels" 2>&1 | tail -15) > gpurun_out/r02_tile_tests.log 2>&1
for wl in cfg2 cfg5; do
  for ty in 8 16; do
    WS_TILE_TY=$ty timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02k_${wl}_ty$ty.json
    WS_TILE_TY=$ty WS_MARCH_DEBUG=1 timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02k_${wl}_ty${ty}_dbg.json
  done
  timeout 300 python bench.py --workload $wl --variant 2 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02k_${wl}_march.json
done
cat gpurun_out/r02_tile_tests.log
for f in gpurun_out/r02k_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
