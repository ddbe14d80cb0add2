#!/bin/bash
# A/B of developer builds (generic library, G = 32 only) on Go2 4096 at the long protocol; two passes to see the run-to-run spread
for pass in 1 2; do
for L in "$@"; do
  GSR_NO_SPEC=1 GSR_LIB=$PWD/genesis_b200/csrc/$L timeout 600 python bench.py --workload go2 --steps 1000 --warmup 5 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('go2', '$L', 'value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
done
