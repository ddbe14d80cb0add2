#!/bin/bash
# launch-geometry sweep of the lane-group kernel: warps per CTA (phase barriers couple the warps of a CTA) and diagnostics
for W in 1 2 4 8; do
  GSR_GROUP_SIZE=32 GSR_WARPS=$W timeout 300 python tools/sweep_one.py go2 4096 2>&1 | grep "##" | sed "s/^/warps=$W /"
done
GSR_GROUP_SIZE=32 GSR_WARPS=2 GSR_DIAGNOSTICS=0 timeout 300 python tools/sweep_one.py go2 4096 2>&1 | grep "##" | sed "s/^/warps=2 /"
GSR_GROUP_SIZE=16 GSR_WARPS=2 timeout 300 python tools/sweep_one.py go2 4096 2>&1 | grep "##" | sed "s/^/warps=2 /"
GSR_GROUP_SIZE=16 GSR_WARPS=4 timeout 300 python tools/sweep_one.py go2 4096 2>&1 | grep "##" | sed "s/^/warps=4 /"
for W in 1 2 4; do
  GSR_GROUP_SIZE=8 GSR_WARPS=$W timeout 300 python tools/sweep_one.py franka 30000 2>&1 | grep "##" | sed "s/^/warps=$W /"
done
GSR_GROUP_SIZE=8 GSR_WARPS=4 GSR_DIAGNOSTICS=0 timeout 300 python tools/sweep_one.py franka 30000 2>&1 | grep "##" | sed "s/^/warps=4 /"
GSR_GROUP_SIZE=16 GSR_WARPS=4 GSR_DIAGNOSTICS=0 timeout 300 python tools/sweep_one.py franka 30000 2>&1 | grep "##" | sed "s/^/warps=4 /"
