#!/bin/bash
# round 2, GPU run Z: the drop-in classes' device sort() and consumers; wall time of sort() through the drop-in at C4 size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dropin.py -m gpu -x -q > gpurun_out/pytest_r02z.log 2>&1; echo "dropin rc $?"; tail -15 gpurun_out/pytest_r02z.log
timeout 900 python profiles/perf_dropin_sort.py > gpurun_out/perf_dropin_sort_r02z.txt 2>&1; tail -8 gpurun_out/perf_dropin_sort_r02z.txt
