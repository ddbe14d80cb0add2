#!/bin/bash
# lazy kinematics (GSR_LAZY_FK=1) against eager with the model-specialised copies, long protocol, twice
for pass in 1 2; do
for Z in 0 1; do
for W in franka go2; do
  GSR_LAZY_FK=$Z timeout 600 python bench.py --workload $W --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W lazy=$Z', 'value %.2fM'%(d['value']/1e6), 'e2e %.2fM'%(d['e2e']['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
done
done
