#!/bin/bash
# final numbers of the marching kernels: bench lines of BASELINE configs 2-5, ncu launch list + full capture (cfg3, cfg4)
mkdir -p gpurun_out
for wl in cfg2 cfg3 cfg4 cfg5; do
  steps=20; case $wl in cfg2|cfg5) steps=200;; esac
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/final_bench_$wl.json
  python -c "import json;d=json.load(open('gpurun_out/final_bench_$wl.json'));print('$wl',round(d['value'],2),d['config']['kernels'],round(d['roofline']['frac'],3),round(d['roofline']['whole_step_frac'],3))"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/march_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --nx 512 --ny 512 --nz 512 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:kMarch -s 6 -c 2 --csv --log-file gpurun_out/march_traffic_cfg3_1024.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
for wl in cfg3 cfg4; do
  n=512; [ $wl = cfg4 ] && n=384
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:kMarch -s 8 -c 2 -f -o gpurun_out/prof_march_$wl python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --nx $n --ny $n --nz $n > gpurun_out/ncu_$wl.log 2>&1
  echo "ncu $wl rc=$?"
done
