#!/bin/bash
# round 2, GPU run AT (2 GPUs): the bench under torchrun with the build that ships
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02at_2gpu.json 2> gpurun_out/bench_r02at_2gpu.err
echo "rc $?"; tail -2 gpurun_out/bench_r02at_2gpu.err; head -c 330 gpurun_out/bench_r02at_2gpu.json
