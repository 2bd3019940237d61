#!/bin/bash
# round 2, GPU run AJ: tensor-core pass, second mapping, stores without divergence
mkdir -p gpurun_out
for cfg in C3 C2 C4; do timeout 300 python profiles/perf_ab.py $cfg tile=4 2>&1 | tail -1; done > gpurun_out/perf_ab_r02aj.txt 2>&1
cat gpurun_out/perf_ab_r02aj.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_mma_r02aj python profiles/prof_target.py 3 > gpurun_out/ncu_interp_mma_r02aj.log 2>&1
