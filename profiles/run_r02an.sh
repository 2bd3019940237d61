#!/bin/bash
# round 2, GPU run AN: staged prefetch also in the generic second location kernel (Nest, Mesh); kernel A back to the L2 prefetch
mkdir -p gpurun_out
for cfg in C3 C3nest C3mesh C4; do timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1; done > gpurun_out/perf_ab_r02an.txt 2>&1
cat gpurun_out/perf_ab_r02an.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r02an.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_r02an.log
