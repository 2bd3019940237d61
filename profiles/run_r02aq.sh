#!/bin/bash
# round 2, GPU run AQ: second trellis location kernel on runs of 4 consecutive tiles per warp (node header and staged records reused)
mkdir -p gpurun_out
for cfg in C3 C2 C5; do timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1; done > gpurun_out/perf_ab_r02aq.txt 2>&1
cat gpurun_out/perf_ab_r02aq.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_reference_regressions.py -m gpu -x -q > gpurun_out/pytest_r02aq.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_r02aq.log
