#!/bin/bash
# 2 B200: halo exchange by the library's own kernels over peer memory against ncclSend / ncclRecv: bitwise parity tests of both
# transports (multi-process: CUDA IPC; host driver: one process, peer access), then strong / weak scaling A/B
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_multigpu_nccl.py -q -x 2>&1 | tail -15) > gpurun_out/r02_p2p_tests.log 2>&1
(timeout 900 python -m pytest tests/test_host_layer.py -q -m gpu -k "several_gpus or spatial_slabs" 2>&1 | tail -8) >> gpurun_out/r02_p2p_tests.log 2>&1
cat gpurun_out/r02_p2p_tests.log
n=2
for tr in 1 0; do
  for wl in cfg3 cfg2; do
    WS_P2P=$tr timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$tr bench.py --gpus $n --workload $wl --strong --steps 40 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02p_strong_${wl}_n${n}_p2p$tr.json
  done
  WS_P2P=$tr timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$tr bench.py --gpus $n --steps 20 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/r02p_weak_northstar_n${n}_p2p$tr.json
done
for f in gpurun_out/r02p_*_n${n}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]; print("%-44s %.2f Gpt/s  %s  halo=%s ms/step %.3f  parity %s" % (sys.argv[1][11:], d["value"], d["config"]["kernels"], d["config"].get("halo"), d["ms_per_step"], d.get("parity_multi")))
except Exception as e:
    print(sys.argv[1], "parse error", e, open(sys.argv[1]).read()[-400:])
PY
done
