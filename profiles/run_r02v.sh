#!/bin/bash
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled > gpurun_out/thp_r02v.txt 2>&1
timeout 600 python profiles/perf_call_sizes.py > gpurun_out/perf_call_sizes_r02v.log 2>&1
cat gpurun_out/thp_r02v.txt; tail -12 gpurun_out/perf_call_sizes_r02v.log
