#!/bin/bash
# 2-D step outside the two half-step kernels: steps per CUDA graph, no graph at all
mkdir -p gpurun_out
for g in 8 32 128; do
  WS_GRAPH_STEPS=$g timeout 300 python bench.py --workload cfg2 --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02m_cfg2_g$g.json
done
WS_NO_GRAPH=1 timeout 300 python bench.py --workload cfg2 --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02m_cfg2_nograph.json
WS_GRAPH_STEPS=32 timeout 300 python bench.py --workload cfg5 --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02m_cfg5_g32.json
for f in gpurun_out/r02m_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  whole %.3f" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
