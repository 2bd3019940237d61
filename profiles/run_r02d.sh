#!/bin/bash
# round 2, GPU run D: the drop-in against brille's own tests, ncu of the cooperative location kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_dropin.py tests/test_sort_oracle.py tests/test_gpu_round2.py -m gpu -q > gpurun_out/pytest_r02d.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_r02d.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trellis_in_node_coop -s 1 -c 1 -f -o gpurun_out/ncu_coop_r02d python profiles/prof_target.py 3 > gpurun_out/ncu_coop_r02d.log 2>&1
tail -15 gpurun_out/pytest_r02d.log
