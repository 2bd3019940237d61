#!/bin/bash
# round 2, GPU run O: ncu captures for the roofline traffic of C4 / C2 / C3 and the launch list of the bench command
mkdir -p gpurun_out
PROF_CONFIG=C4 timeout 600 ncu --set full --clock-control none -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_c4_r02o python profiles/prof_target.py 3 5e5 > gpurun_out/ncu_c4_r02o.log 2>&1
PROF_CONFIG=C2 timeout 600 ncu --set full --clock-control none -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_c2_r02o python profiles/prof_target.py 3 1e6 > gpurun_out/ncu_c2_r02o.log 2>&1
PROF_CONFIG=C3nest timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_locate_in_node -s 1 -c 1 -f -o gpurun_out/ncu_locate_nest_r02o python profiles/prof_target.py 3 > gpurun_out/ncu_nest_r02o.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02o.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r02o.log 2>&1
ls -la gpurun_out/*r02o*
