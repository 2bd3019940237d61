#!/bin/bash
# round 2, GPU run AD (final state): whole GPU suite, smoke, bench + reference arm, launch list of the bench command, ncu --set full of
# the three hot kernels (traffic of the interpolation kernel with the padded records; local traffic / scoreboard of the location kernels)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02ad.log 2>&1; echo "suite rc $?"; tail -4 gpurun_out/pytest_r02ad.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02ad.log 2>&1; tail -2 gpurun_out/smoke_r02ad.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02ad_1gpu.json 2> gpurun_out/bench_r02ad_1gpu.err; echo "bench rc $?"; head -c 600 gpurun_out/bench_r02ad_1gpu.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02ad_reference.json 2> gpurun_out/bench_r02ad_reference.err; echo "reference rc $?"; head -c 300 gpurun_out/bench_r02ad_reference.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_cell_tma -s 1 -c 1 -f -o gpurun_out/ncu_interp_r02ad python profiles/prof_target.py 3 > gpurun_out/ncu_interp_r02ad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_locate<|k_trellis_in_node_coop" -s 2 -c 2 -f -o gpurun_out/ncu_locate_r02ad python profiles/prof_target.py 3 > gpurun_out/ncu_locate_r02ad.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02ad.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r02ad.log 2>&1
ls -la gpurun_out/*r02ad*
