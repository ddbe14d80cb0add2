#!/bin/bash
# Round-2 session A: parity suite, smoke, bench at the driver's protocol (--steps 20 --warmup 5) and a long run, reference arm.
set -u
TAG=${1:-r2a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
( timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
tail -15 $O/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v ptxas | tail -5 > $O/smoke.log )
cat $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; tail -1 $O/bench_driver.json | cut -c1-400; tail -3 $O/bench_driver.err
timeout 900 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > $O/bench_long.json 2> $O/bench_long.err; tail -1 $O/bench_long.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2>&1; tail -1 $O/bench_reference.json | cut -c1-300
