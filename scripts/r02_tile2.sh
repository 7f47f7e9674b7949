#!/bin/bash
# full GPU suite + smoke with the tile kernels and the fused acquisition launch; cfg2 / cfg5 lines; ncu launch list + full capture
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r02_gpu_tests2.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8) > gpurun_out/r02_smoke2.log 2>&1
for wl in cfg2 cfg5; do
  timeout 300 python bench.py --workload $wl --steps 200 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02l_${wl}.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_cfg2_tile.csv python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu --no-others > /dev/null 2>&1
for wl in cfg2 cfg5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:kTile2D -s 8 -c 2 -f -o gpurun_out/r02_ncu_tile_$wl python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu_t_$wl.log 2>&1
  echo "ncu $wl rc=$?"
done
cat gpurun_out/r02_gpu_tests2.log gpurun_out/r02_smoke2.log
for f in gpurun_out/r02l_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.2f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f launches %s" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"], d.get("gpu_launches")))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
