#!/bin/bash
# programmatic dependent launch on the 2-D path: parity tests, then A/B
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_wavefield_objects.py -q -x -m gpu -k "tile2d or default_kernels or wavefield or step_scaling or golden or ci_cases" 2>&1 | tail -6) > gpurun_out/r02_pdl_tests.log 2>&1
cat gpurun_out/r02_pdl_tests.log
for p in 1 0 1 0; do
  for wl in cfg2 cfg5; do
    WS_PDL=$p timeout 300 python bench.py --workload $wl --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02n_${wl}_pdl$p.json
  done
done
for f in gpurun_out/r02n_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-36s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  whole %.3f finite %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["whole_step_frac"], d["config"]["finite"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
