#!/bin/bash
# round 2, GPU run Q (2 GPUs): the bench under torchrun on 2 ranks (all legs incl. the powder all-reduce), reference arm
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02q_2gpu.json 2> gpurun_out/bench_r02q_2gpu.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices" > gpurun_out/pytest_r02q.log 2>&1
tail -3 gpurun_out/bench_r02q_2gpu.err; head -c 400 gpurun_out/bench_r02q_2gpu.json; tail -2 gpurun_out/pytest_r02q.log
