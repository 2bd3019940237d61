#!/bin/bash
# round 2, GPU run AH: ncu of the tensor-core variant of the cell kernel (first mapping: 8 points x 4 slots per warp step)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_mma_r02ah python profiles/prof_target.py 3 > gpurun_out/ncu_interp_mma_r02ah.log 2>&1
ls -la gpurun_out/ncu_interp_mma_r02ah.ncu-rep
