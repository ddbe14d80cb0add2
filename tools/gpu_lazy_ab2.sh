#!/bin/bash
# one shared kinematics routine (new specialised copies) against the copies with the kinematics inlined at both call sites
# (kept aside as libgsrigid.keep_inl_*.so), lazy and eager, long protocol
run() {  # label, workload, env...
  local label=$1 W=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $W --steps 1000 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W $label', 'value %.2fM'%(d['value']/1e6), 'e2e %.2fM'%(d['e2e']['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'], d['config'].get('kinematics','')[:5], d['config'].get('specialised_library'))"
}
for pass in 1 2; do
for W in franka go2; do
  run "shared-routine lazy(default)" $W GSR_X=0
  run "shared-routine eager" $W GSR_LAZY_FK=0
  run "inlined lazy" $W GSR_LIB=$PWD/genesis_b200/csrc/libgsrigid.keep_inl_$W.so GSR_LAZY_FK=1
done
done
