#!/bin/bash
# round 2, GPU run N: scan with 8 buckets per thread (fine regrouping bins again): whole GPU suite + timings
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02n.log 2>&1
for c in C3 C3nest C3mesh C4; do timeout 300 python profiles/perf_ab.py $c > gpurun_out/perf_${c}_r02n.log 2>&1; done
timeout 300 python profiles/perf_ab.py C2 nq=1e6 > gpurun_out/perf_C2_r02n.log 2>&1
tail -3 gpurun_out/pytest_r02n.log; cat gpurun_out/perf_C*_r02n.log | cut -c1-250
