#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift; local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --variant 3 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02c_${wl}_${name}.json
}
for wl in cfg2 cfg5; do
  run dbg $wl WS_MARCH_DEBUG=1 WS_TMA_STAGES=3
  run dbg6 $wl WS_MARCH_DEBUG=1 WS_TMA_STAGES=6
done
WS_TMA_STAGES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kTma -s 6 -c 2 -o gpurun_out/r02_ncu_tma_cfg2 python bench.py --workload cfg2 --variant 3 --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
for f in gpurun_out/r02c_cfg*_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-36s %.1f Gpt/s  %s  ms %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-300:])
PY
done
tail -3 gpurun_out/ncu_cfg2.log
