#!/bin/bash
# round 2, GPU run R: regrouping-bin size of the Nest / Mesh location with the tight candidate lists
mkdir -p gpurun_out
for c in C3nest C3mesh; do timeout 300 python profiles/perf_ab.py $c bin_points=30 bin_points=300 bin_points=30 bin_points=300 > gpurun_out/perf_${c}_r02r.log 2>&1; done
cat gpurun_out/perf_C*_r02r.log | cut -c1-250
