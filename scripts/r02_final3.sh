#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py 2>gpurun_out/r02_final3_bench.err | tail -1 > gpurun_out/r02_final3_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kTma -s 8 -c 1 -f -o gpurun_out/r02_ncu_tma_cfg4_v2 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-others --nx 384 --ny 384 --nz 384 > gpurun_out/ncu_cfg4_v2.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final3_bench.json").read())
print({k:(v if k not in ("config","roofline","cpu_baseline") else "...") for k,v in d.items()})
print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("traffic_source","kernel_timing")})
for k,v in (d["config"].get("others") or {}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!="workload"})
print("cpu", d.get("cpu_baseline"))
PY
