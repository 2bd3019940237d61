#!/bin/bash
# round 2, GPU run AI: tensor-core compute pass, second mapping (8 slots x 8 points per warp step, a lane finishes one slot for two points)
mkdir -p gpurun_out
for cfg in C3 C2 C4; do timeout 300 python profiles/perf_ab.py $cfg tile=4 tile=2 2>&1 | tail -2; done > gpurun_out/perf_ab_r02ai.txt 2>&1
cat gpurun_out/perf_ab_r02ai.txt
timeout 300 python profiles/perf_consumer.py > gpurun_out/perf_consumer_r02ai.txt 2>&1; tail -6 gpurun_out/perf_consumer_r02ai.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_consumer.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r02ai.log 2>&1; echo "pytest rc $?"; tail -12 gpurun_out/pytest_r02ai.log
