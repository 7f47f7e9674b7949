#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r02_gpu_tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8) > gpurun_out/r02_smoke.log 2>&1
for wl in cfg2 cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02e_${wl}.json
done
WS_NO_FAST_A=1 timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02e_cfg4_nofastA.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-others 2>&1 | tail -1 > gpurun_out/r02e_northstar.json
cat gpurun_out/r02_gpu_tests.log gpurun_out/r02_smoke.log
for f in gpurun_out/r02e_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.1f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
    if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"])
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
