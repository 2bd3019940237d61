#!/bin/bash
# round 2, GPU run M: Nest / Mesh fast path (sphere prefilter) with the coarser sort bins: parity + timing
mkdir -p gpurun_out
timeout 2000 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_consumer.py -m gpu -q -x -k "nest or mesh or Nest or Mesh or c4 or C4 or golden or shared or two_kernel" > gpurun_out/pytest_r02m.log 2>&1
for c in C3nest C3mesh C4; do timeout 300 python profiles/perf_ab.py $c > gpurun_out/perf_${c}_r02m.log 2>&1; done
tail -3 gpurun_out/pytest_r02m.log; cat gpurun_out/perf_C*_r02m.log | cut -c1-250
