#!/bin/bash
# usage: exp.sh name:ENV=VAL,ENV2=VAL:benchargs ...   — ncu DRAM/L2 traffic + duration of the tiled kernels at 1024^3
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum
for spec in "$@"; do
  IFS=':' read -r name envs bargs <<< "$spec"
  envs=${envs//,/ }
  env A=1 $envs timeout 900 ncu --metrics $M --clock-control none -k regex:kFast -s 8 -c 4 --csv --log-file gpurun_out/exp_$name.csv python bench.py --steps 1 --warmup 3 --no-cpu $bargs > gpurun_out/exp_$name.log 2>&1
  python - "$name" <<'PY'
import csv,sys
name=sys.argv[1]
N=1024.0**3
try:
    rows=[r for r in csv.reader(open('gpurun_out/exp_%s.csv'%name)) if len(r)>10 and r[0].isdigit()]
except Exception as e:
    print(name,'FAILED',e); sys.exit(0)
d={}
for r in rows: d.setdefault((r[0],r[4]),{})[r[12]]=float(r[14].replace(',',''))
import json
tot={}
for k,v in d.items():
    kn='vel' if 'Vel' in k[1] else 'str'
    ms=v['gpu__time_duration.sum']/1e6; rd=v['dram__bytes_read.sum']; wr=v['dram__bytes_write.sum']
    print('%-14s %s %-40s %7.3f ms  rd %5.1f B/pt  wr %5.1f B/pt  dram %5.0f GB/s  l2rd %5.1f B/pt' % (name,kn,k[1][-40:],ms,rd/N,wr/N,(rd+wr)/ms/1e6,v['lts__t_sectors_srcunit_tex_op_read.sum']*32/N))
    t=tot.setdefault(kn,{'ms':0,'dram_read_bytes':0,'dram_write_bytes':0,'launches':0})
    t['ms']+=ms; t['dram_read_bytes']+=rd; t['dram_write_bytes']+=wr; t['launches']+=1
json.dump({'grid':'1024^3','source':'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one time step','per_half_step':tot}, open('gpurun_out/traffic_%s.json'%name,'w'), indent=1)
PY
done
