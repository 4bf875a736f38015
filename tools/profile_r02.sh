#!/bin/bash
# Round-2 profiling pass (run under gpurun): launch list of the eager bench step, full captures of the heavy kernels.
# The .ncu-rep files are converted to csv on the box and removed (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out/prof
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-graph"
$CMD > gpurun_out/prof/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/prof/launches.csv $CMD > gpurun_out/prof/ncu_launch.log 2>&1
cap() {   # kernel regex, skip, count, command...
  k=$1; s=$2; c=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/prof/$k "$@" > gpurun_out/prof/ncu_$k.log 2>&1
  ncu -i gpurun_out/prof/$k.ncu-rep --page raw --csv > gpurun_out/prof/$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof/$k.ncu-rep --page source --csv > gpurun_out/prof/$k.src.csv 2>/dev/null
  rm -f gpurun_out/prof/$k.ncu-rep
}
cap knn_tcp_scan 8 1 $CMD
cap knn_tc_rerank_kernel 8 1 $CMD
cap knn_xyz_kernel 4 1 $CMD
cap edge_gather_reduce 12 3 $CMD
cap edge_bwd_scatter 12 3 $CMD
cap gemm_tc_kernel 16 4 $CMD
python tools/full_step_once.py 2 > gpurun_out/prof/plain_full.log 2>&1 || exit 1
cap gf_forward 2 1 python tools/full_step_once.py 3
cap op_forward 2 1 python tools/full_step_once.py 3
cap op_bwd_main 1 1 python tools/full_step_once.py 3
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/prof/launches_full.csv python tools/full_step_once.py 3 > gpurun_out/prof/ncu_launch_full.log 2>&1
du -sh gpurun_out/prof
