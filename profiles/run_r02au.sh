#!/bin/bash
# round 2, GPU run AU (the build that ships): compute-sanitizer over every kernel family (small sizes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_target.py > gpurun_out/sanitize_memcheck_r02au.log 2>&1; echo "memcheck rc $?" >> gpurun_out/sanitize_memcheck_r02au.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_target.py > gpurun_out/sanitize_racecheck_r02au.log 2>&1; echo "racecheck rc $?" >> gpurun_out/sanitize_racecheck_r02au.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/sanitize_target.py > gpurun_out/sanitize_synccheck_r02au.log 2>&1; echo "synccheck rc $?" >> gpurun_out/sanitize_synccheck_r02au.log
tail -6 gpurun_out/sanitize_memcheck_r02au.log; tail -6 gpurun_out/sanitize_racecheck_r02au.log; tail -4 gpurun_out/sanitize_synccheck_r02au.log
