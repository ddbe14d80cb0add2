#!/bin/bash
# developer sweep: Go2 4096 with 16 lanes per environment (two environments per warp) and different on-chip row capacities
for cfg in "32 0" "16 16" "16 24" "16 32" "16 12"; do
  set -- $cfg
  export GSR_GROUP_SIZE=$1
  if [ "$2" = "0" ]; then unset GSR_SMEM_ROWS; else export GSR_SMEM_ROWS=$2; fi
  python bench.py --workload go2 --steps 1000 --warmup 300 --no-cpu-baseline --no-extra-legs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('G=$1 rows=$2', 'value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2fM'%(d['e2e']['value']/1e6), d['config'].get('launch'))"
done
