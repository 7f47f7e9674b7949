#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/r02_final_scale_n8.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final_scale_n8.json").read())
print(round(d["value"],2), d["ms_per_step"], d["config"]["halo"], d.get("parity_multi"), d["config"]["finite"], d["clocks"])
PY
