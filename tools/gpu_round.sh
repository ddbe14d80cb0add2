#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full capture of the step kernel. Outputs under gpurun_out/.
set -u
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
tail -5 $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v ptxas | tail -5 > $O/smoke.log )
cat $O/smoke.log
timeout 600 python bench.py --steps 1000 --warmup 100 > $O/bench_franka.json 2> $O/bench_franka.err; tail -1 $O/bench_franka.json
timeout 600 python bench.py --workload go2 --steps 1000 --warmup 100 > $O/bench_go2.json 2> $O/bench_go2.err; tail -1 $O/bench_go2.json
timeout 300 python bench.py --impl reference --steps 200 --warmup 10 > $O/bench_reference.json 2>&1; tail -1 $O/bench_reference.json
for W in franka go2; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $O/launches_$W.csv \
     python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_launch_$W.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 2 -f -o $O/prof_$W \
     python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_full_$W.log 2>&1
done
ls -la $O
