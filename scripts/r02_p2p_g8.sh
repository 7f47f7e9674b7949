#!/bin/bash
# 8 B200: the peer-memory halo exchange at 4 ranks (bitwise parity tests) and at 8 ranks (weak scaling of the north-star, strong scaling of
# configs 3 and 2) next to ncclSend / ncclRecv
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_nccl.py -q -x -k "4-peer or 4-nccl" 2>&1 | tail -6) > gpurun_out/r02_p2p_tests_w4.log 2>&1
cat gpurun_out/r02_p2p_tests_w4.log
n=8
for tr in 1 0; do
  WS_P2P=$tr timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$tr bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02p_weak_northstar_n${n}_p2p$tr.json
  for wl in cfg3 cfg2; do
    WS_P2P=$tr timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$tr bench.py --gpus $n --workload $wl --strong --steps 40 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02p_strong_${wl}_n${n}_p2p$tr.json
  done
done
for f in gpurun_out/r02p_*_n${n}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("%-44s %.2f Gpt/s  %s  halo=%s ms/step %.3f  parity %s" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], d["config"].get("halo"), d["ms_per_step"], d.get("parity_multi")))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-400:])
PY
done
