#!/bin/bash
# round 2, GPU run X (4 GPUs): the bench under torchrun on 4 ranks
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r02x_4gpu.json 2> gpurun_out/bench_r02x_4gpu.err
tail -3 gpurun_out/bench_r02x_4gpu.err; head -c 300 gpurun_out/bench_r02x_4gpu.json
