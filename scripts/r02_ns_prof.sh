#!/bin/bash
# north-star with the producer-warpgroup kernels in one launch per half-step: DRAM traffic at 1024^3, launch list of the headline command,
# full ncu capture at 512^3
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:kFast -s 6 -c 2 --csv --log-file gpurun_out/exp_r02u.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/exp_r02u.log 2>&1
python - <<'PY'
import csv,json
N=1024.0**3
rows=[r for r in csv.reader(open('gpurun_out/exp_r02u.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((r[0],r[4]),{})[r[12]]=float(r[14].replace(',',''))
tot={}
for k,v in d.items():
    kn='vel' if 'Vel' in k[1] else 'str'
    ms=v['gpu__time_duration.sum']/1e6; rd=v['dram__bytes_read.sum']; wr=v['dram__bytes_write.sum']
    print('%s %-40s %7.3f ms  rd %5.1f B/pt  wr %5.1f B/pt  dram %5.0f GB/s  l2rd %5.1f B/pt' % (kn,k[1][-40:],ms,rd/N,wr/N,(rd+wr)/ms/1e6,v['lts__t_sectors_srcunit_tex_op_read.sum']*32/N))
    t=tot.setdefault(kn,{'ms':0,'dram_read_bytes':0,'dram_write_bytes':0,'launches':0})
    t['ms']+=ms; t['dram_read_bytes']+=rd; t['dram_write_bytes']+=wr; t['launches']+=1
json.dump({'grid':'1024^3','source':'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one time step (round 2: producer-warpgroup kernels, one launch per half-step)','per_half_step':tot}, open('gpurun_out/traffic_r02u.json','w'), indent=1)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_northstar_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kFast -s 6 -c 2 -f -o gpurun_out/r02_ncu_fast_512_v2 python bench.py --steps 2 --warmup 3 --no-cpu --no-others --nx 512 --ny 512 --nz 512 > gpurun_out/ncu_fast_v2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_ncu_fast_512_v2.ncu-rep
