#!/bin/bash
# north-star, one launch per half-step (WS_FAST_FLAGS=0): chunk length, hints, repeatability (20 steps)
mkdir -p gpurun_out; rm -f gpurun_out/r02t_*
run() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02t_$TAG.json; }
TAG=f2 run WS_FAST_FLAGS=2
TAG=f0 run WS_FAST_FLAGS=0
TAG=f0b run WS_FAST_FLAGS=0
TAG=f0ch32 run WS_FAST_FLAGS=0 WS_FAST_CHUNK=32
TAG=f0ch48 run WS_FAST_FLAGS=0 WS_FAST_CHUNK=48
TAG=f0ch96 run WS_FAST_FLAGS=0 WS_FAST_CHUNK=96
TAG=f0ch128 run WS_FAST_FLAGS=0 WS_FAST_CHUNK=128
TAG=f1 run WS_FAST_FLAGS=1
for f in gpurun_out/r02t_*.json; do python - "$f" <<'PY'
import json,sys
for ln in open(sys.argv[1]).read().strip().splitlines():
    try:
        d=json.loads(ln)
        r=d["roofline"]; print("%-28s %.2f Gpt/s  ms/step %.3f  kernels %.3f/%.3f  frac %.3f whole %.3f e2e %.2f" % (sys.argv[1][11:], d["value"], d["ms_per_step"], r["ms_first"], r["ms_second"], r["frac"], r["whole_step_frac"], d["e2e"]["value"]))
    except Exception as e:
        print(sys.argv[1], "parse error", e, ln[-300:])
PY
done
