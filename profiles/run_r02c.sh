#!/bin/bash
# round 2, GPU run C: rest of the GPU suite, occupancy variants of the cooperative location kernel, ncu of it, sort launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02c.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_r02c.log
for v in 5 6; do BRILLE_B200_LIB=$PWD/profiles/variants/lib_coop$v.so timeout 300 python profiles/perf_ab.py C3 coop_locate=1 > gpurun_out/perf_coop${v}_r02c.log 2>&1; done
timeout 300 python profiles/perf_ab.py C3 coop_locate=1 > gpurun_out/perf_coop4_r02c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trellis_in_node_coop -s 3 -c 1 -f -o gpurun_out/ncu_coop_r02c python profiles/prof_target.py 3 > gpurun_out/ncu_coop_r02c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort_r02c.csv python profiles/perf_sort.py > gpurun_out/perf_sort_ncu_r02c.log 2>&1
tail -5 gpurun_out/pytest_r02c.log; cat gpurun_out/perf_coop*_r02c.log
