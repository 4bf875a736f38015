#!/bin/bash
# ncu captures (source-level) of the xyz tensor-core kNN: scan and re-rank kernels.  Run under gpurun.
mkdir -p gpurun_out/px
CMD="python tools/xyz_tc_once.py 2"
$CMD > gpurun_out/px/plain.log 2>&1 || exit 1
cap() {   # kernel regex, skip, count
  k=$1; s=$2; c=$3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/px/$k $CMD > gpurun_out/px/ncu_$k.log 2>&1
  ncu -i gpurun_out/px/$k.ncu-rep --page raw --csv > gpurun_out/px/$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/px/$k.ncu-rep --page source --csv > gpurun_out/px/$k.src.csv 2>/dev/null
  rm -f gpurun_out/px/$k.ncu-rep
}
cap knn_tcp_scan 3 1
cap knn_tc_rerank_kernel 6 1
