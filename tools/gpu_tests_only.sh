#!/bin/bash
# parity suite only (default kernels)
O=gpurun_out/${1:-tests}
mkdir -p $O
( timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log )
grep -E "^FAILED|^ERROR|passed|failed" $O/pytest_gpu.log | head -40
