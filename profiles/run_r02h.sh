#!/bin/bash
# round 2, GPU run H: powder-average consumer tests + full bench (new e2e legs, powder leg, per-config block)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_consumer.py -m gpu -q -x > gpurun_out/pytest_r02h.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
tail -5 gpurun_out/pytest_r02h.log; tail -3 gpurun_out/bench_r02h.err; head -c 300 gpurun_out/bench_r02h.json
