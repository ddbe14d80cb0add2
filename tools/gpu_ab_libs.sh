#!/bin/bash
# A/B of alternative builds of libgsrigid (GSR_LIB): franka long protocol
for L in "$@"; do
  GSR_LIB=$PWD/genesis_b200/csrc/$L timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', 'value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
