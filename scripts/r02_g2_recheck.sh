#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_multigpu_nccl.py -q -x 2>&1 | tail -5) > gpurun_out/r02_g2_recheck.log 2>&1
(timeout 900 python -m pytest tests/test_host_layer.py -q -m gpu -k "several_gpus or spatial_slabs" 2>&1 | tail -3) >> gpurun_out/r02_g2_recheck.log 2>&1
cat gpurun_out/r02_g2_recheck.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02w_weak_northstar_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w_weak_northstar_n2.json").read())
print(round(d["value"],2), d["ms_per_step"], d["config"]["halo"], d.get("parity_multi"), d["config"]["finite"])
PY
