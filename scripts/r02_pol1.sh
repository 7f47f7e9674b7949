#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift
  env "$@" timeout 400 python bench.py --edge-policy 1 --steps 10 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02g_pol1_${name}.json
}
run def X=1
run l4 WS_MARCH_LANES=4
run st3 WS_TMA_STAGES=3
run march X=1 WS_NO_TMA_MARCH=1
for f in gpurun_out/r02g_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
