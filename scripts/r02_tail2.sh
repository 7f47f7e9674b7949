#!/bin/bash
# 2-D path as the product runs it (graphs): acquisition tail and programmatic dependent launch A/B; the failing host test with output
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_host_layer.py -q -x -m gpu -k "wavefield_operators_on_gpu or compensation_on_gpu or product_driver_on_gpu" 2>&1 | tail -40) > gpurun_out/r02_tail_tests2.log 2>&1
cat gpurun_out/r02_tail_tests2.log
rm -f gpurun_out/r02q_*
for rep in 1 2; do
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  for wl in cfg2 cfg5; do
    WS_TILE_TAIL=$1 WS_PDL=$2 timeout 300 python bench.py --workload $wl --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02q_${wl}_tail$1_pdl$2.json
  done
done
done
for f in gpurun_out/r02q_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-36s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  whole %.3f finite %s launches %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["whole_step_frac"], d["config"]["finite"], d.get("gpu_launches")))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
