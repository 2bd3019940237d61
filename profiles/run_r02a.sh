#!/bin/bash
# round 2, GPU run A: new parity tests, smoke, bench with the per-config block, ncu of the location kernels, host topology
mkdir -p gpurun_out
( nvidia-smi topo -m; echo; grep -i allowed /proc/self/status; ls /sys/devices/system/node; nproc; free -g; lscpu | head -30; which numactl ) > gpurun_out/topo_r02a.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_r02a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02a.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_locate -s 4 -c 2 -f -o gpurun_out/ncu_locate_r02a python profiles/prof_target.py 3 > gpurun_out/ncu_locate_r02a.log 2>&1
tail -3 gpurun_out/pytest_r02a.log; cat gpurun_out/smoke_r02a.log | tail -3; head -c 600 gpurun_out/bench_r02a.json
