#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "marching and tma" 2>&1 | tail -5) > gpurun_out/r02_tma_tests4.log 2>&1
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --variant 3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02d_${wl}_${name}.json
}
for wl in cfg2 cfg5; do
  run dbg $wl WS_MARCH_DEBUG=1 WS_TMA_STAGES=3
  for st in 2 3 4 6; do run st$st $wl WS_TMA_STAGES=$st; done
  run ch32 $wl WS_TMA_CHUNK=32
  run nl2 $wl WS_MARCH_LANES=2
done
for wl in cfg3 cfg4; do
  run def $wl X=1
  run st4 $wl WS_TMA_STAGES=4
done
cat gpurun_out/r02_tma_tests4.log
for f in gpurun_out/r02d_cfg*_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.1f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
