#!/bin/bash
# round 2, GPU run E: interpolation kernel, unrolled-corner variant A/B + full ncu profile with source
mkdir -p gpurun_out
timeout 600 python profiles/perf_ab.py C3 interp_variant=0 interp_variant=1 interp_variant=0 interp_variant=1 > gpurun_out/perf_ab_r02e.log 2>&1
timeout 600 python profiles/perf_ab.py C4 interp_variant=0 interp_variant=1 >> gpurun_out/perf_ab_r02e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_r02e python profiles/prof_target.py 3 > gpurun_out/ncu_interp_r02e.log 2>&1
cat gpurun_out/perf_ab_r02e.log
