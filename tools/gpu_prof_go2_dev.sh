#!/bin/bash
# ncu --set full of one steady-state Go2 launch with a developer build (GSR_LIB), per-line attribution restricted to the narrow phase
O=gpurun_out/${1:-r2u}
LIB=$PWD/genesis_b200/csrc/${2:-libgsrigid.dev_g32.so}
mkdir -p $O
GSR_NO_SPEC=1 GSR_LIB=$LIB timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 345 -c 1 -f -o $O/prof_go2_dev \
   python bench.py --workload go2 --steps 30 --warmup 5 --no-cpu-baseline --no-extra-legs > $O/ncu_full_go2_dev.log 2>&1
export GSR_PROF_LIB=$LIB
python tools/ncu_funcs.py $O/prof_go2_dev.ncu-rep k_stepILi32ELb1ELb1 > $O/prof_go2_dev.funcs.txt 2>&1
GSR_LINES_FILTER=gsr_collide.cuh python tools/ncu_lines.py $O/prof_go2_dev.ncu-rep k_stepILi32ELb1ELb1 70 > $O/prof_go2_dev.lines.txt 2>&1
rm -f $O/prof_go2_dev.ncu-rep
head -24 $O/prof_go2_dev.funcs.txt; tail -72 $O/prof_go2_dev.lines.txt | cut -c1-200
