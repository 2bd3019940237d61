#!/bin/bash
# round 2, GPU run AB: 128-byte per-point records (one line per bulk copy) and item blocks of 16 / 4 against the default build
mkdir -p gpurun_out
for cfg in C3 C2; do
  echo "== $cfg default"; timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1
  for v in rec128 ib16 ib4; do echo "== $cfg $v"; BRILLE_B200_LIB=$PWD/profiles/variants/lib_$v.so timeout 300 python profiles/perf_ab.py $cfg 2>&1 | tail -1; done
done > gpurun_out/perf_ab_r02ab.txt 2>&1
cat gpurun_out/perf_ab_r02ab.txt
