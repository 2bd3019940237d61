#!/bin/bash
# round 2, GPU run AP: ncu with source of the second trellis location kernel after the staged prefetch
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trellis_in_node_coop -s 1 -c 1 -f -o gpurun_out/ncu_locate_b_r02ap python profiles/prof_target.py 3 > gpurun_out/ncu_locate_b_r02ap.log 2>&1
