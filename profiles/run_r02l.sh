#!/bin/bash
# round 2, GPU run L: Nest / Mesh fast path with bounding-box prefilter: whole GPU suite, timing, full bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02l.log 2>&1
for c in C3nest C3mesh C4; do timeout 300 python profiles/perf_ab.py $c > gpurun_out/perf_${c}_r02l.log 2>&1; done
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err
tail -4 gpurun_out/pytest_r02l.log; cat gpurun_out/perf_C*_r02l.log | cut -c1-250; tail -2 gpurun_out/bench_r02l.err
