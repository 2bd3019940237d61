#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair_match -s 1 -c 1 -f -o gpurun_out/ncu_match_c3_r02u python profiles/prof_sort_target.py > gpurun_out/ncu_match_c3_r02u.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort_r02u.csv python profiles/prof_sort_target.py > /dev/null 2>&1
grep -E "k_pair" gpurun_out/launches_sort_r02u.csv | cut -d, -f5,11-15 | head
