#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -5) > gpurun_out/r02_final4_tests.log 2>&1
cat gpurun_out/r02_final4_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2)
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02_final4_cfg3.json
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02_final4_cfg4.json
for f in gpurun_out/r02_final4_cfg*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); r=d["roofline"]
print("%-28s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
PY
done
