#!/bin/bash
# full GPU suite, smoke, the default bench line (as the driver runs it) and the reference arm
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/r02_final_tests.log 2>&1
cat gpurun_out/r02_final_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r02_final_smoke.log 2>&1
cat gpurun_out/r02_final_smoke.log
timeout 900 python bench.py 2>gpurun_out/r02_final_bench.err | tail -1 > gpurun_out/r02_final_bench.json
for wl in cfg2 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02_final_$wl.json
done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final_bench.json").read())
print({k:(v if k not in ("config","roofline","cpu_baseline") else "...") for k,v in d.items()})
print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("traffic_source",)})
print("others", json.dumps(d["config"].get("others"))[:1500])
print("cpu", d.get("cpu_baseline"))
for wl in ("cfg2","cfg5"):
    e=json.loads(open("gpurun_out/r02_final_%s.json"%wl).read()); r=e["roofline"]
    print(wl, round(e["value"],2), r["ms_first"], r["ms_second"], round(r["frac"],3), round(r["whole_step_frac"],3))
PY
