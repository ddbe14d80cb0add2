#!/bin/bash
# developer A/B: Franka 30000 with alternative builds (GSR_LIB): launch geometry, per-phase timeline, long bench
for L in "$@"; do
  export GSR_LIB=$PWD/genesis_b200/csrc/$L
  echo "== $L"
  timeout 300 python tools/step_trace.py franka 2>&1 | grep -E "kernel span|env step duration|phase|resident" | head -9
  python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2fM'%(d['e2e']['value']/1e6))"
done
