#!/bin/bash
# round 2, GPU run AR: run length of the second trellis location kernel (2 / 4 / 8 / 16 consecutive tiles per warp)
mkdir -p gpurun_out
( echo "== run 4 (default)"; timeout 300 python profiles/perf_ab.py C3 2>&1 | tail -1
for r in 2 8 16; do echo "== run $r"; BRILLE_B200_LIB=$PWD/profiles/variants/lib_run$r.so timeout 300 python profiles/perf_ab.py C3 2>&1 | tail -1; done ) > gpurun_out/perf_ab_r02ar.txt 2>&1
cat gpurun_out/perf_ab_r02ar.txt
