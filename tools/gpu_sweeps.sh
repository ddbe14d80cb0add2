#!/bin/bash
# Launch-geometry sweeps behind the defaults of gsr_host.h / gsr_api.cu (results: profiles/README.md "Sweeps").
#   bash tools/gpu_sweeps.sh franka|go2|small
# Tuning knobs ride environment variables read at scene build: GSR_GROUP_SIZE (lanes per environment, 1 = lane kernel),
# GSR_WARPS (warps per CTA), GSR_SMEM_ROWS (constraint rows kept on chip), GSR_DIAGNOSTICS, GSR_REORDER_PERIOD.
line() { tail -1 | cut -c1-120; }
case "${1:-franka}" in
franka)
  for G in 8 16; do for W in 1 2 4; do
    GSR_GROUP_SIZE=$G GSR_WARPS=$W timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | line | sed "s/^/franka G=$G warps=$W /"
  done; done
  for R in 24 16 8; do
    GSR_SMEM_ROWS=$R timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | line | sed "s/^/franka rows=$R /"
  done
  GSR_REORDER_PERIOD=0 timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | line | sed "s/^/franka no-reorder /"
  GSR_DIAGNOSTICS=0 timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | line | sed "s/^/franka no-diagnostics /"
  ;;
go2)
  for R in 24 32 40 64; do for W in 6 7 8; do
    GSR_SMEM_ROWS=$R GSR_WARPS=$W timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | line | sed "s/^/go2 rows=$R warps=$W /"
  done; done
  GSR_REORDER_PERIOD=0 timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | line | sed "s/^/go2 no-reorder /"
  GSR_GROUP_SIZE=16 timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | line | sed "s/^/go2 G=16 /"
  ;;
small)
  python tools/sweep_small.py 2>&1 | grep "##"
  ;;
esac
