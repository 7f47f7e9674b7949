#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02v_*
run() { env "$@" timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02v_$TAG.json; }
TAG=def run A=1
TAG=ch32 run WS_TMA_CHUNK=32
TAG=ch48 run WS_TMA_CHUNK=48
TAG=ch96 run WS_TMA_CHUNK=96
TAG=fch48 run WS_FAST_CHUNK=48
TAG=fch96 run WS_FAST_CHUNK=96
TAG=st3 run WS_TMA_STAGES=3
for f in gpurun_out/r02v_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-28s %.2f Gpt/s  ms/step %.3f  kernels %.3f/%.3f  frac %.3f whole %.3f" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
