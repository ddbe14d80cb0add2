#!/bin/bash
# steady-state ncu captures of the Go2 step kernel (after 300 steps) for two CTA shapes
O=gpurun_out/go2prof
mkdir -p $O
for W in 8 1; do
  GSR_WARPS=$W timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 300 -c 1 -f -o $O/prof_go2_w$W \
     python bench.py --workload go2 --steps 310 --warmup 5 --no-cpu-baseline > $O/ncu_w$W.log 2>&1
  tail -2 $O/ncu_w$W.log | cut -c1-300
done
for W in 8 4 1; do
  GSR_WARPS=$W timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
done
