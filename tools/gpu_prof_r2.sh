#!/bin/bash
# Round-2 profiling session: ncu launch lists (--metrics gpu__time_duration.sum) and one --set full capture (with source
# counters) of the step kernel per workload; per-function / per-line attributions computed on the box (reports exceed 64 MiB).
set -u
TAG=${1:-r2p}
O=gpurun_out/$TAG
mkdir -p $O
for W in franka go2; do
  S=40; [ $W = go2 ] && S=340
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 60 --csv --log-file $O/launches_$W.csv \
     python bench.py --workload $W --steps $((S + 30)) --warmup 5 --no-cpu-baseline --no-extra-legs > $O/ncu_launch_$W.log 2>&1
  # the full capture skips the internal settling steps of bench.py (32 Franka / 320 Go2) + the warm-up: a steady-state launch
  F=45; [ $W = go2 ] && F=335
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s $F -c 1 -f -o $O/prof_$W \
     python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline --no-extra-legs > $O/ncu_full_$W.log 2>&1
done
for R in franka go2; do
  K=k_stepILi8ELb0ELb0; [ $R = go2 ] && K=k_stepILi32ELb1ELb1
  export GSR_PROF_LIB=$(python tools/which_lib.py $R 2>/dev/null | tail -1)   # the copy that ran (model-specialised for the bench scenes)
  echo "$R: $GSR_PROF_LIB" >> $O/prof_libs.txt
  ncu -i $O/prof_$R.ncu-rep --page raw --csv > $O/prof_$R.raw.csv 2>/dev/null
  python tools/ncu_funcs.py $O/prof_$R.ncu-rep $K > $O/prof_$R.funcs.txt 2>&1
  python tools/ncu_lines.py $O/prof_$R.ncu-rep $K 60 > $O/prof_$R.lines.txt 2>&1
  rm -f $O/prof_$R.ncu-rep
done
( timeout 300 python tools/step_trace.py go2 2>&1 | grep -v Warn > $O/trace_go2.txt; timeout 300 python tools/step_trace.py franka 2>&1 | grep -v -i warn > $O/trace_franka.txt )
ls -la $O
