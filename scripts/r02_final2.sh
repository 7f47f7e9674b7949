#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/r02_final2_tests.log 2>&1
cat gpurun_out/r02_final2_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r02_final2_smoke.log 2>&1
cat gpurun_out/r02_final2_smoke.log
timeout 900 python bench.py 2>gpurun_out/r02_final2_bench.err | tail -1 > gpurun_out/r02_final2_bench.json
timeout 300 python bench.py --edge-policy 1 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02_final2_pol1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final2_bench.json").read())
print({k:(v if k not in ("config","roofline","cpu_baseline") else "...") for k,v in d.items()})
print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("traffic_source","kernel_timing")})
for k,v in (d["config"].get("others") or {}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
print("cpu", d.get("cpu_baseline"))
e=json.loads(open("gpurun_out/r02_final2_pol1.json").read()); r=e["roofline"]
print("pol1", round(e["value"],2), r["ms_first"], r["ms_second"], round(r["frac"],3), round(r["whole_step_frac"],3))
PY
