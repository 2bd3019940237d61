#!/bin/bash
# round 2, GPU run J: powder consumer with the privatised histogram kernel: tests, bench leg, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_consumer.py -m gpu -q -x -k powder > gpurun_out/pytest_r02j.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_powder_r02j.csv python - > gpurun_out/powder_ncu_r02j.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import brille_b200
from brille_b200 import host, workloads as W, _bridge
wl = W.c3_p63mmc(host.get())
g = brille_b200.accelerate(wl.grid)
rng = np.random.default_rng(5)
g.set_structure_factor(rng.normal(size=4) + 1j * rng.normal(size=4), positions=rng.uniform(0, 1, (4, 3)), q_transform=np.asarray(_bridge.flatten_bz(wl.bz)["to_xyz"]).reshape(3, 3))
for _ in range(2):
    h, c = g.ir_powder_sweep((0.1, 10.0), 200, (0.0, 55.0), 400, 50000, seed=7, weight=1)
print(h.sum(), c[0])
PY
tail -3 gpurun_out/pytest_r02j.log; tail -2 gpurun_out/bench_r02j.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r02j.json'))
print(d['consumer']['powder_average'])
PY
