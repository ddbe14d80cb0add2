#!/bin/bash
# Short closing session after a kernel change: parity tests (default + lane kernel), smoke, the three bench lines, and the
# Franka launch list / full capture / phase trace (the Go2 and lane captures of tools/gpu_round.sh stay valid).
set -u
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log ); tail -3 $O/pytest_gpu.log
( GSR_GROUP_SIZE=1 timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_lane.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_lane.log ); tail -3 $O/pytest_gpu_lane.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep smoke > $O/smoke.log ); cat $O/smoke.log | cut -c1-160
timeout 600 python bench.py > $O/bench_franka.json 2> $O/bench_franka.err; tail -1 $O/bench_franka.json | cut -c1-200
timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline > $O/bench_go2.json 2> $O/bench_go2.err; tail -1 $O/bench_go2.json | cut -c1-200
timeout 600 python bench.py --workload grasp --no-cpu-baseline > $O/bench_grasp.json 2> $O/bench_grasp.err; tail -1 $O/bench_grasp.json | cut -c1-200
timeout 300 python tools/step_trace.py franka 2>&1 | grep -v -i warn > $O/trace_franka.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $O/launches_franka.csv \
   python bench.py --workload franka --steps 50 --warmup 5 --no-cpu-baseline > $O/ncu_launch_franka.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 2 -f -o $O/prof_franka \
   python bench.py --workload franka --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_full_franka.log 2>&1
ncu -i $O/prof_franka.ncu-rep --page raw --csv > $O/prof_franka.raw.csv 2>/dev/null
python tools/ncu_funcs.py $O/prof_franka.ncu-rep k_stepILi8ELb0 > $O/prof_franka.funcs.txt 2>&1
python tools/ncu_lines.py $O/prof_franka.ncu-rep k_stepILi8ELb0 40 > $O/prof_franka.lines.txt 2>&1
rm -f $O/prof_franka.ncu-rep
ls $O
