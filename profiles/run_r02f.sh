#!/bin/bash
# round 2, GPU run F: pipelined cell kernel with prefetched task draws + staged item descriptors vs the previous build
mkdir -p gpurun_out
for i in 1 2; do
BRILLE_B200_LIB=$PWD/profiles/variants/lib_head.so timeout 300 python profiles/perf_ab.py C3 > gpurun_out/perf_head_r02f_$i.log 2>&1
timeout 300 python profiles/perf_ab.py C3 interp_variant=0 interp_variant=1 > gpurun_out/perf_new_r02f_$i.log 2>&1
done
BRILLE_B200_LIB=$PWD/profiles/variants/lib_head.so timeout 300 python profiles/perf_ab.py C4 > gpurun_out/perf_head_r02f_c4.log 2>&1
timeout 300 python profiles/perf_ab.py C4 > gpurun_out/perf_new_r02f_c4.log 2>&1
BRILLE_B200_LIB=$PWD/profiles/variants/lib_head.so timeout 300 python profiles/perf_ab.py C2 nq=1e6 > gpurun_out/perf_head_r02f_c2.log 2>&1
timeout 300 python profiles/perf_ab.py C2 nq=1e6 > gpurun_out/perf_new_r02f_c2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or cell_kernels or full_size or degenerate" > gpurun_out/pytest_r02f.log 2>&1
cat gpurun_out/perf_*_r02f_*.log; tail -3 gpurun_out/pytest_r02f.log
