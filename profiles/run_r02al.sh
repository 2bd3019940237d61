#!/bin/bash
# round 2, GPU run AL: ncu --set full with source of the two trellis location kernels as shipped
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_locate -s 1 -c 1 -f -o gpurun_out/ncu_locate_a_r02al python profiles/prof_target.py 3 > gpurun_out/ncu_locate_a_r02al.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trellis_in_node_coop -s 1 -c 1 -f -o gpurun_out/ncu_locate_b_r02al python profiles/prof_target.py 3 > gpurun_out/ncu_locate_b_r02al.log 2>&1
ls -la gpurun_out/*r02al*
