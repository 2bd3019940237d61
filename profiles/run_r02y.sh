#!/bin/bash
# round 2, GPU run Y: the reference's C++ regression cases restated as GPU parity tests, then the whole GPU suite and smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_regressions.py -m gpu -x -q > gpurun_out/pytest_r02y_regr.log 2>&1; echo "regr rc $?"; tail -15 gpurun_out/pytest_r02y_regr.log
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_regressions.py > gpurun_out/pytest_r02y.log 2>&1; echo "suite rc $?"; tail -5 gpurun_out/pytest_r02y.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02y.log 2>&1; tail -2 gpurun_out/smoke_r02y.log
