#!/bin/bash
# round 2, GPU run AE: point records packed in shared memory (96-byte stride) with the 128-byte stride in global memory
mkdir -p gpurun_out
for cfg in C3 C2 C4; do timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1; done > gpurun_out/perf_ab_r02ae.txt 2>&1
cat gpurun_out/perf_ab_r02ae.txt
timeout 300 python profiles/perf_consumer.py > gpurun_out/perf_consumer_r02ae.txt 2>&1; tail -12 gpurun_out/perf_consumer_r02ae.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_r02ae python profiles/prof_target.py 3 > gpurun_out/ncu_interp_r02ae.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_sf_r02ae python profiles/prof_target.py 3 1e7 sf > gpurun_out/ncu_interp_sf_r02ae.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_consumer.py -m gpu -x -q > gpurun_out/pytest_r02ae.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_r02ae.log
