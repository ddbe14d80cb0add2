#!/bin/bash
# A/B of alternative builds of libgsrigid (GSR_LIB) at the long protocol.  Usage: gpu_ab_libs.sh <workload> <lib.so> [<lib.so> ...]
W=$1; shift
for L in "$@"; do
  GSR_LIB=$PWD/genesis_b200/csrc/$L timeout 600 python bench.py --workload $W --steps 1000 --warmup 300 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W', '$L', 'value %.2fM'%(d['value']/1e6), 'kernel_ms %.4f'%d['roofline']['kernel_ms'])"
done
