#!/bin/bash
# re-run of the profiling half of tools/gpu_r2final.sh
bash tools/gpu_prof_r2.sh ${1:-r2z} > gpurun_out/${1:-r2z}/prof.log 2>&1
cat gpurun_out/${1:-r2z}/prof_libs.txt; head -12 gpurun_out/${1:-r2z}/prof_franka.funcs.txt
