#!/bin/bash
# Closing session of round 2 on one GPU: smoke, the bench lines (driver protocol, long protocol, Go2, grasp, reference arm), then the
# ncu launch lists + full captures + per-function / per-line attributions + phase traces of tools/gpu_prof_r2.sh.
set -u
TAG=${1:-r2z}
O=gpurun_out/$TAG
mkdir -p $O
( timeout 300 python __graft_entry__.py smoke 2>&1 | grep -i smoke > $O/smoke.log ); cut -c1-200 $O/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; tail -1 $O/bench_driver.json | cut -c1-300
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; tail -1 $O/bench_reference.json | cut -c1-300
timeout 900 python bench.py --steps 1000 --warmup 100 --no-extra-legs > $O/bench_franka.json 2> $O/bench_franka.err; tail -1 $O/bench_franka.json | cut -c1-200
timeout 900 python bench.py --workload go2 --steps 1000 --warmup 300 > $O/bench_go2.json 2> $O/bench_go2.err; tail -1 $O/bench_go2.json | cut -c1-200
timeout 900 python bench.py --workload grasp --steps 500 --warmup 100 --no-cpu-baseline > $O/bench_grasp.json 2> $O/bench_grasp.err; tail -1 $O/bench_grasp.json | cut -c1-200
bash tools/gpu_prof_r2.sh $TAG > $O/prof.log 2>&1
ls $O
