#!/bin/bash
# round 2, GPU run AA: per-point records of the pipelined cell kernel by 16-byte asynchronous copies (LDGSTS, sector granular)
# against one 96-byte bulk copy per point (which fetches whole 128-byte lines): A/B on C3, C2, C4; parity subset on the new build
mkdir -p gpurun_out
for cfg in C3 C2 C4; do
  echo "== $cfg gather (default build)"; timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1
  echo "== $cfg bulk";   BRILLE_B200_LIB=$PWD/profiles/variants/lib_bulk.so timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1
done > gpurun_out/perf_ab_r02aa.txt 2>&1
cat gpurun_out/perf_ab_r02aa.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r02aa.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_r02aa.log
