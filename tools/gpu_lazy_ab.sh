#!/bin/bash
# lazy vs eager kinematics on one box: parity suite in both modes, then the long bench lines
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 )
( GSR_EAGER_FK=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 )
for e in 0 1 0 1; do
  for w in franka go2; do
    GSR_EAGER_FK=$e python bench.py --workload $w --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eager=$e $w', 'value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.2fM'%(d['e2e']['value']/1e6))"
  done
done
