#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fast or default_kernels or baseline" 2>&1 | tail -12) > gpurun_out/r02_ns_tests.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-others --no-cpu 2>&1 | tail -1 > gpurun_out/r02f_northstar.json
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02f_cfg4.json
cat gpurun_out/r02_ns_tests.log
for f in gpurun_out/r02f_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
