#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_long.py -q -s -k drift 2>&1 | tail -12) > gpurun_out/r02_drift.log 2>&1
(timeout 600 python scripts/stability_cfg3.py 1024 2>&1 | tail -14) > gpurun_out/r02_stability_cfg3_1024.txt 2>&1
# ncu: launch list of the headline command, and full captures of the dominant kernels of configs 3, 4, 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_northstar.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kTma -s 6 -c 2 -o gpurun_out/r02_ncu_tma_cfg3 python bench.py --workload cfg3 --nx 512 --ny 512 --nz 512 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kTma|kFastVel" -s 12 -c 3 -o gpurun_out/r02_ncu_tma_cfg4 python bench.py --workload cfg4 --nx 384 --ny 384 --nz 384 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kMarch -s 6 -c 2 -o gpurun_out/r02_ncu_march_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c5.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"kTma|kFastVel" -s 12 -c 3 --csv --log-file gpurun_out/r02_traffic_cfg4_768.csv python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_t4.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:kTma -s 6 -c 2 --csv --log-file gpurun_out/r02_traffic_cfg3_1024.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_t3.log 2>&1
cat gpurun_out/r02_drift.log gpurun_out/r02_stability_cfg3_1024.txt; ls -la gpurun_out/*.ncu-rep
