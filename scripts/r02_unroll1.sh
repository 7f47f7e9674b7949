#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02C_*
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "visco" 2>&1 | tail -3)
for wl in cfg4 cfg5; do for i in 1 2; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02C_$wl.json; done; done
for f in gpurun_out/r02C_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-28s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  frac %.3f whole %.3f finite %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"], d["config"]["finite"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
