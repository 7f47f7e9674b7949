#!/bin/bash
# GPU check of the TMA marching kernels: parity tests, then BASELINE configs 2-5 with sweeps of ring depth / chunk length
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "marching" 2>&1 | tail -25) > gpurun_out/r02_tma_tests.log 2>&1
run() { # name, env..., workload
  local name=$1; shift; local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --variant 3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02_${wl}_${name}.json
}
for wl in cfg2 cfg3 cfg4 cfg5; do
  run def $wl X=1
  for st in 2 3 4 6; do run st$st $wl WS_TMA_STAGES=$st; done
  for ch in 32 128; do run ch$ch $wl WS_TMA_CHUNK=$ch; done
done
run l4 cfg4 WS_MARCH_LANES=4
run l1 cfg3 WS_MARCH_LANES=1
cat gpurun_out/r02_tma_tests.log
for f in gpurun_out/r02_cfg*_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.1f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
