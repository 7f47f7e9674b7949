#!/bin/bash
# A/B on one box: tiled kernels of HEAD (gpurun_ab_old.so) against the working tree (order-reducing edges in the tiled kernels)
mkdir -p gpurun_out
for i in 1 2; do
WAVESIM_LIB=$PWD/gpurun_ab_old.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02j_old_$i.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02j_new_$i.json
done
for f in gpurun_out/r02j_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
