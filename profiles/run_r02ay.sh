#!/bin/bash
# round 2, GPU run AY: the cooperative second location kernel at 5 / 6 CTAs per SM (96 / 80 registers, some spills) against 4
mkdir -p gpurun_out
( echo "== 4 CTAs / SM (default)"; timeout 300 python profiles/perf_ab.py C3 2>&1 | tail -1
for v in 5 6; do echo "== $v CTAs / SM"; BRILLE_B200_LIB=$PWD/profiles/variants/lib_coop$v.so timeout 300 python profiles/perf_ab.py C3 2>&1 | tail -1; done
echo "== 4 CTAs / SM (default) again"; timeout 300 python profiles/perf_ab.py C3 2>&1 | tail -1 ) > gpurun_out/perf_ab_r02ay.txt 2>&1
cat gpurun_out/perf_ab_r02ay.txt
