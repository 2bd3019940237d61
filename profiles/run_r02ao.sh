#!/bin/bash
# round 2, GPU run AO: first location kernel with the search trips of moveinto compacted through a per-warp queue
mkdir -p gpurun_out
( for cfg in C3 C2; do
  echo "== $cfg default build (no register cap)"; timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1
  for v in loc3 loc4; do echo "== $cfg $v"; BRILLE_B200_LIB=$PWD/profiles/variants/lib_$v.so timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1; done
done ) > gpurun_out/perf_ab_r02ao.txt 2>&1
cat gpurun_out/perf_ab_r02ao.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_reference_regressions.py -m gpu -x -q > gpurun_out/pytest_r02ao.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_r02ao.log
