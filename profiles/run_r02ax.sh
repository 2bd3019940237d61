#!/bin/bash
# round 2, GPU run AX: ncu with source of the second location kernel on a Nest (C3 lattice)
mkdir -p gpurun_out
PROF_CONFIG=C3nest timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_locate_in_node -s 1 -c 1 -f -o gpurun_out/ncu_locate_nest_r02ax python profiles/prof_target.py 3 > gpurun_out/ncu_locate_nest_r02ax.log 2>&1
