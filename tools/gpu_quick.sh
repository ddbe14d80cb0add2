#!/bin/bash
# quick check after a kernel change: parity suite (default kernels) + steady-state bench lines of both workloads
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
timeout 600 python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-180
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-180
