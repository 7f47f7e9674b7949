#!/bin/bash
# tiled 3-D elastic kernels: planes per unrolled trip (UNR 1 default, 2, 4) under the 152-register budget
mkdir -p gpurun_out; rm -f gpurun_out/r02y_*
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02y_unr1.json
for u in 2 4; do
  WAVESIM_LIB=$PWD/gpurun_ab_unr$u.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02y_unr$u.json
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02y_unr1.json
for f in gpurun_out/r02y_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-28s %.2f Gpt/s  ms/step %.3f  kernels %.3f/%.3f  frac %.3f whole %.3f finite %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"], d["config"]["finite"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
