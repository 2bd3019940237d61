#!/bin/bash
# round 2, GPU run AC (8 GPUs): the bench under torchrun on 8 ranks, as the driver launches it
mkdir -p gpurun_out
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02ac_8gpu.json 2> gpurun_out/bench_r02ac_8gpu.err
echo "rc $?"; tail -3 gpurun_out/bench_r02ac_8gpu.err; head -c 400 gpurun_out/bench_r02ac_8gpu.json
