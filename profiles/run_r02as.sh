#!/bin/bash
# round 2, GPU run AS (the build that ships, after the location-kernel changes): whole GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02as.log 2>&1; echo "suite rc $?"; tail -4 gpurun_out/pytest_r02as.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02as.log 2>&1; tail -2 gpurun_out/smoke_r02as.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02as_1gpu.json 2> gpurun_out/bench_r02as_1gpu.err; echo "bench rc $?"; head -c 400 gpurun_out/bench_r02as_1gpu.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trellis_in_node_coop -s 1 -c 1 -f -o gpurun_out/ncu_locate_b_r02as python profiles/prof_target.py 3 > gpurun_out/ncu_locate_b_r02as.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02as.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r02as.log 2>&1
