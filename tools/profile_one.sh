#!/bin/bash
# ncu --set full + source page of ONE kernel of the eager bench step: tools/profile_one.sh <kernel regex> <skip> [out name]
k=$1; s=$2; o=${3:-$1}
mkdir -p gpurun_out/p1
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-graph"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c 1 -o gpurun_out/p1/$o $CMD > gpurun_out/p1/ncu_$o.log 2>&1
ncu -i gpurun_out/p1/$o.ncu-rep --page raw --csv > gpurun_out/p1/$o.raw.csv 2>/dev/null
ncu -i gpurun_out/p1/$o.ncu-rep --page source --csv > gpurun_out/p1/$o.src.csv 2>/dev/null
rm -f gpurun_out/p1/$o.ncu-rep
