#!/bin/bash
# round 2, GPU run G: staged item descriptors alone vs previous build; sort() with the staged cost kernel; bench with the new e2e legs
mkdir -p gpurun_out
for i in 1 2; do
BRILLE_B200_LIB=$PWD/profiles/variants/lib_head.so timeout 300 python profiles/perf_ab.py C3 > gpurun_out/perf_head_r02g_$i.log 2>&1
timeout 300 python profiles/perf_ab.py C3 > gpurun_out/perf_new_r02g_$i.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sort_oracle.py -m gpu -q -x -k "sort or solver" > gpurun_out/pytest_r02g.log 2>&1
timeout 600 python profiles/perf_sort.py > gpurun_out/perf_sort_r02g.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err
cat gpurun_out/perf_*_r02g_*.log | cut -c1-220; tail -3 gpurun_out/pytest_r02g.log; tail -2 gpurun_out/perf_sort_r02g.log; tail -3 gpurun_out/bench_r02g.err; head -c 300 gpurun_out/bench_r02g.json
