#!/bin/bash
# round 2, GPU run P: tighter candidate lists (plane test) for the Nest / Mesh fast path: parity + timing; then the profile captures of run O
mkdir -p gpurun_out
timeout 2000 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_consumer.py tests/test_dropin.py -m gpu -q -x -k "nest or mesh or Nest or Mesh or c4 or C4 or golden or shared or two_kernel or dropin_launches" > gpurun_out/pytest_r02p.log 2>&1
for c in C3nest C3mesh C4; do timeout 300 python profiles/perf_ab.py $c > gpurun_out/perf_${c}_r02p.log 2>&1; done
tail -3 gpurun_out/pytest_r02p.log; cat gpurun_out/perf_C*_r02p.log | cut -c1-250
bash profiles/run_r02o.sh
