#!/bin/bash
mkdir -p gpurun_out
for nl in 2 1; do
(WS_MARCH_LANES=$nl timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "marching and tma and 2D" 2>&1 | tail -5) > gpurun_out/r02_tma_tests_nl$nl.log 2>&1
done
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --variant 3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02b_${wl}_${name}.json
}
for wl in cfg2 cfg5; do
  for nl in 4 2 1; do for st in 2 3 4 6; do run nl${nl}st$st $wl WS_MARCH_LANES=$nl WS_TMA_STAGES=$st; done; done
  for nl in 2 1; do for ch in 32 128; do run nl${nl}ch$ch $wl WS_MARCH_LANES=$nl WS_TMA_CHUNK=$ch; done; done
done
cat gpurun_out/r02_tma_tests_nl*.log
for f in gpurun_out/r02b_cfg*_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.1f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
