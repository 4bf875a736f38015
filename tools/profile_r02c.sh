#!/bin/bash
# Closing profiling pass of round 2 (run under gpurun): launch list of the eager bench step, full captures of the three scan /
# re-rank kernels after the Hilbert-order / xyz-on-tensor-cores changes.  .ncu-rep files are converted to csv and removed.
mkdir -p gpurun_out/prof
CMD="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-graph"
$CMD > gpurun_out/prof/plain.log 2>&1 || exit 1
[ -z "$SKIP_LAUNCH_LIST" ] && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/prof/launches.csv $CMD > gpurun_out/prof/ncu_launch.log 2>&1
cap() {   # output name, kernel regex, skip, count, command...
  o=$1; k=$2; s=$3; c=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c $c -o gpurun_out/prof/$o "$@" > gpurun_out/prof/ncu_$o.log 2>&1
  ncu -i gpurun_out/prof/$o.ncu-rep --page raw --csv > gpurun_out/prof/$o.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof/$o.ncu-rep --page source --csv > gpurun_out/prof/$o.src.csv 2>/dev/null
  rm -f gpurun_out/prof/$o.ncu-rep
}
# -k matches the function name without template arguments; per step the launches are scan <3>, <64>, <64> and
# re-rank <3,0>, <3,1>, <64,0>, <64,1>, <64,0>, <64,1>: the skip counts pick the instance
if [ -z "$SKIP_CAPTURES" ]; then
cap scan64 knn_tcp_scan_kernel 8 1 $CMD
cap scan3 knn_tcp_scan_kernel 6 1 $CMD
cap rerank64 knn_tc_rerank_kernel 14 1 $CMD
cap rerank3 knn_tc_rerank_kernel 12 1 $CMD
fi
du -sh gpurun_out/prof
