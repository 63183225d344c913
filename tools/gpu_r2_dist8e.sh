#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multigpu_8e.log 2>&1
tail -4 gpurun_out/r2_pytest_multigpu_8e.log
{
for cfg in c5 c3; do
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 8 $cfg 3
  echo "head off:"; GSP_CHOL_FUSED_HEAD=0 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 8 $cfg 3
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=8 python tools/gpu_dist.py 8 $cfg 3
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 4 $cfg 3
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 1 $cfg 3
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_dist_sweep_8gpu_e.log
GSP_PROF_TIMELINE=1 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python - > gpurun_out/r2_timeline_c5_8gpu_e.txt 2>&1 <<'PY'
import sys, os, numpy as np
sys.path.insert(0, '.')
import gsp_b200 as gsp
lib = gsp.Library(devices=list(range(8)))
dims, nd, kind = ((256, 128), 500, 1)
N = dims[0] * dims[1]
A = np.zeros((3, 3)); A[0, 0] = A[1, 1] = 1 / 20.0
rng = np.random.default_rng(3)
dinds = np.sort(rng.choice(N, nd, replace=False)); z1 = rng.standard_normal(nd) * 0.5
dom = (gsp._lib.make_grid_domain(dims, [0.0, 0.0], [1.0, 1.0]), None)
gsp.LUPlan(lib, [(kind, 1.0, A)], dom, dinds + 1, z1, 0.0).close()
lib.profile_enable(True)
p = gsp.LUPlan(lib, [(kind, 1.0, A)], dom, dinds + 1, z1, 0.0)
print(lib.profile_read(), p.times())
PY
tail -1 gpurun_out/r2_timeline_c5_8gpu_e.txt | cut -c1-600
