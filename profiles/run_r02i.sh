#!/bin/bash
# round 2, GPU run I: certified plane filter in the trellis tetrahedron search (per-lane second kernel) vs the cooperative kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "not sort and not nest and not mesh and not c4" > gpurun_out/pytest_r02i.log 2>&1
timeout 300 python profiles/perf_ab.py C3 coop_locate=1 coop_locate=0 coop_locate=1 coop_locate=0 > gpurun_out/perf_ab_r02i.log 2>&1
timeout 300 python profiles/perf_ab.py C2 coop_locate=1 coop_locate=0 nq=1e6 >> gpurun_out/perf_ab_r02i.log 2>&1
timeout 300 python profiles/perf_ab.py C1 coop_locate=1 coop_locate=0 >> gpurun_out/perf_ab_r02i.log 2>&1
tail -4 gpurun_out/pytest_r02i.log; cut -c1-260 gpurun_out/perf_ab_r02i.log
