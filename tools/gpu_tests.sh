#!/bin/bash
# parity suite on the GPU for the default kernel choice and for the lane kernel
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 )
( GSR_GROUP_SIZE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 )
