#!/bin/bash
# round 2, GPU run AK (the build that ships): whole GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_r02ak.log 2>&1; echo "suite rc $?"; tail -4 gpurun_out/pytest_r02ak.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02ak.log 2>&1; tail -2 gpurun_out/smoke_r02ak.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02ak_1gpu.json 2> gpurun_out/bench_r02ak_1gpu.err; echo "bench rc $?"; head -c 400 gpurun_out/bench_r02ak_1gpu.json; echo
