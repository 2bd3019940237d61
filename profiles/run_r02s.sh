#!/bin/bash
# round 2, GPU run S: occupancy of the assignment solver (40 / 32 registers) against the default build
mkdir -p gpurun_out
timeout 300 python profiles/perf_sort.py > gpurun_out/perf_sort_r02s_default.log 2>&1
for v in 6 8; do BRILLE_B200_LIB=$PWD/profiles/variants/lib_match$v.so timeout 300 python profiles/perf_sort.py > gpurun_out/perf_sort_r02s_$v.log 2>&1; done
for f in gpurun_out/perf_sort_r02s_*.log; do echo $f; cut -c1-110 $f; done
