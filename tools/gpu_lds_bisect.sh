#!/bin/bash
# which file's workspace accesses may be addressed as shared memory (LDS / STS) without changing results?
# for every developer build: the parity suite (failures listed) and the franka / go2 numbers
O=gpurun_out/${1:-lds}
shift
mkdir -p $O
for L in "$@"; do
  export GSR_LIB=$PWD/genesis_b200/csrc/$L
  ( timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_properties.py > $O/pytest_$L.log 2>&1 )
  echo "== $L: $(grep -E 'passed|failed' $O/pytest_$L.log | tail -1)"; grep -E "^FAILED" $O/pytest_$L.log | head -5
  timeout 600 python bench.py --steps 500 --warmup 100 --no-cpu-baseline --no-extra-legs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   franka value %.2fM kernel_ms %.4f'%(d['value']/1e6, d['roofline']['kernel_ms']))"
done
