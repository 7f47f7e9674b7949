#!/bin/bash
# SASS evidence for profiles/: per kernel of libwavesim_cuda.so the number of TMA loads / stores (UTMALDG / UTMASTG), bulk
# copies (UBLKCP), cp.async (LDGSTS), mbarrier operations (SYNCS) and warp shuffles (SHFL).  usage: sass_summary.sh > profiles/rNN_sass_summary.txt
SO=${1:-wave-simulation_b200/csrc/libwavesim_cuda.so}
echo "cuobjdump -sass $SO: instruction counts per kernel (sm_100a)"
printf "%-110s %8s %8s %8s %8s %8s %8s\n" kernel UTMALDG UTMASTG UBLKCP LDGSTS SYNCS SHFL
cuobjdump -sass "$SO" | awk '
/Function : / { if (name != "") printf "%s %d %d %d %d %d %d\n", name, a, b, c, d, e, f; name=$3; a=b=c=d=e=f=0 }
/UTMALDG/ {a++} /UTMASTG/ {b++} /UBLKCP/ {c++} /LDGSTS/ {d++} /SYNCS/ {e++} /SHFL/ {f++}
END { if (name != "") printf "%s %d %d %d %d %d %d\n", name, a, b, c, d, e, f }' | while read n a b c d e f; do
  printf "%-110s %8d %8d %8d %8d %8d %8d\n" "$(echo $n | c++filt | sed 's/(WsParams.*//; s/void //; s/(anonymous namespace):://' | cut -c1-110)" $a $b $c $d $e $f
done | sort
