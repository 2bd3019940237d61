#!/bin/bash
# round 2, GPU run B: full GPU test suite with the new kernels (cooperative second location kernel, warp-per-pair assignment
# solver), A/B of the location kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_r02b.log
timeout 600 python profiles/perf_ab.py C3 coop_locate=0 coop_locate=1 > gpurun_out/perf_ab_r02b.log 2>&1
timeout 600 python profiles/perf_ab.py C2 coop_locate=0 coop_locate=1 nq=1e6 >> gpurun_out/perf_ab_r02b.log 2>&1
timeout 600 python profiles/perf_sort.py > gpurun_out/perf_sort_r02b.log 2>&1
tail -5 gpurun_out/pytest_r02b.log; cat gpurun_out/perf_ab_r02b.log; tail -5 gpurun_out/perf_sort_r02b.log
