#!/bin/bash
# round 2, GPU run T: serial bidding for few modes: sort tests + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sort_oracle.py -m gpu -q -x -k "sort or solver" > gpurun_out/pytest_r02t.log 2>&1
timeout 300 python profiles/perf_sort.py > gpurun_out/perf_sort_r02t.log 2>&1
tail -2 gpurun_out/pytest_r02t.log; cut -c1-230 gpurun_out/perf_sort_r02t.log
