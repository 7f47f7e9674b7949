#!/bin/bash
# tiled velocity half-step with on-the-fly inverse densities (WS_FAST_RHO=1 default) against the staged arrays (WS_FAST_RHO=0)
mkdir -p gpurun_out; rm -f gpurun_out/r02x_*
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fast or default_kernels or linearity or golden" 2>&1 | tail -4) > gpurun_out/r02_rho_tests.log 2>&1
cat gpurun_out/r02_rho_tests.log
for i in 1 2; do
for r in 1 0; do
  WS_FAST_RHO=$r timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02x_rho$r.json
done
done
WS_FAST_RHO=1 timeout 300 python bench.py --edge-policy 1 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02x_pol1_rho1.json
WS_FAST_RHO=1 timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02x_cfg4_rho1.json
for f in gpurun_out/r02x_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-28s %.2f Gpt/s  ms/step %.3f  kernels %.3f/%.3f  frac %.3f whole %.3f finite %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"], d["config"]["finite"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
