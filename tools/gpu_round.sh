#!/bin/bash
# One GPU session: parity tests (default kernels + lane kernel), smoke, bench lines, per-env phase traces, ncu launch lists and
# full captures of the step kernels.  Outputs under gpurun_out/<tag>/ (summarised into profiles/ by tools/ncu_summary.py).
set -u
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
tail -3 $O/pytest_gpu.log
( GSR_GROUP_SIZE=1 timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_lane.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_lane.log )
tail -3 $O/pytest_gpu_lane.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v ptxas | tail -5 > $O/smoke.log )
cat $O/smoke.log
timeout 600 python bench.py --steps 1000 --warmup 100 > $O/bench_franka.json 2> $O/bench_franka.err; tail -1 $O/bench_franka.json | cut -c1-200
timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 > $O/bench_go2.json 2> $O/bench_go2.err; tail -1 $O/bench_go2.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 200 --warmup 10 > $O/bench_reference.json 2>&1; tail -1 $O/bench_reference.json | cut -c1-200
GSR_GROUP_SIZE=1 timeout 600 python bench.py --steps 300 --warmup 50 --no-cpu-baseline > $O/bench_franka_lane.json 2> $O/bench_franka_lane.err; tail -1 $O/bench_franka_lane.json | cut -c1-200
( timeout 300 python tools/step_trace.py go2 2>&1 | grep -v Warn > $O/trace_go2.txt; timeout 300 python tools/step_trace.py franka 2>&1 | grep -v -i warn > $O/trace_franka.txt )
for W in franka go2; do
  S=20; [ $W = go2 ] && S=320
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 60 --csv --log-file $O/launches_$W.csv \
     python bench.py --workload $W --steps $((S + 30)) --warmup 5 --no-cpu-baseline > $O/ncu_launch_$W.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s $((S / 2)) -c 2 -f -o $O/prof_$W \
     python bench.py --workload $W --steps $((S / 2 + 10)) --warmup 5 --no-cpu-baseline > $O/ncu_full_$W.log 2>&1
done
GSR_GROUP_SIZE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 1 -f -o $O/prof_franka_lane \
     python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_full_franka_lane.log 2>&1
# gpurun copies back at most 64 MiB: keep the csv exports / attributions, drop the raw reports
for R in franka go2 franka_lane; do
  K=k_stepILi8ELb0; [ $R = go2 ] && K=k_stepILi32ELb1;  # the Franka bench scene has no MPR pair, Go2 with self-collision does
   [ $R = franka_lane ] && K=k_step_lane
  ncu -i $O/prof_$R.ncu-rep --page raw --csv > $O/prof_$R.raw.csv 2>/dev/null
  python tools/ncu_funcs.py $O/prof_$R.ncu-rep $K > $O/prof_$R.funcs.txt 2>&1
  python tools/ncu_lines.py $O/prof_$R.ncu-rep $K 40 > $O/prof_$R.lines.txt 2>&1
  rm -f $O/prof_$R.ncu-rep
done
ls -la $O
