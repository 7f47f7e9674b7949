#!/bin/bash
# traffic / time decomposition of the tiled kernels at 1024^3 (ncu with a handful of metrics = few replays)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum
run() { # name, env, bench args
  echo "== $1"
  env $2 timeout 900 ncu --metrics $M --clock-control none -k regex:kFast -s 6 -c 2 --csv --log-file gpurun_out/exp_$1.csv python bench.py --steps 1 --warmup 3 --no-cpu $3 > gpurun_out/exp_$1.log 2>&1
  python - "$1" <<'PY'
import csv,sys
name=sys.argv[1]
rows=[r for r in csv.reader(open('gpurun_out/exp_%s.csv'%name)) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((r[0],r[4][:40]),{})[r[12]]=float(r[14].replace(',',''))
for k,v in d.items():
    print(name,k[1],' '.join('%s=%.4g'%(a.split('__')[-1][:28],b) for a,b in v.items()))
PY
}
run base "A=1" ""
run dbg1 "WS_FAST_DEBUG=1" ""
run nocpml "A=1" "--damping 0"
run nofs "A=1" "--free-surface 0"
run nocpml_nofs "A=1" "--damping 0 --free-surface 0"
