#!/bin/bash
# ncu source-level capture of the re-rank kernel of one xyz kNN call.  Run under gpurun.
mkdir -p gpurun_out/pr
CMD="python tools/xyz_tc_once.py 2"
k=knn_tc_rerank4_kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/pr/$k $CMD > gpurun_out/pr/ncu_$k.log 2>&1
ncu -i gpurun_out/pr/$k.ncu-rep --page raw --csv > gpurun_out/pr/$k.raw.csv 2>/dev/null
ncu -i gpurun_out/pr/$k.ncu-rep --page source --csv > gpurun_out/pr/$k.src.csv 2>/dev/null
rm -f gpurun_out/pr/$k.ncu-rep
