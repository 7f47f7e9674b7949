#!/bin/bash
# acquisition in the tail of the second half-step (2-D tile kernels): parity tests, then A/B
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -6) > gpurun_out/r02_tail_tests.log 2>&1
cat gpurun_out/r02_tail_tests.log
rm -f gpurun_out/r02o_*
for p in 1 0 1 0; do
  for wl in cfg2 cfg5; do
    WS_TILE_TAIL=$p timeout 300 python bench.py --workload $wl --steps 256 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 >> gpurun_out/r02o_${wl}_tail$p.json
  done
done
for f in gpurun_out/r02o_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-36s %.2f Gpt/s  ms/step %.4f  kernels %.3f/%.3f  whole %.3f finite %s launches %s" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["whole_step_frac"], d["config"]["finite"], d.get("gpu_launches")))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
