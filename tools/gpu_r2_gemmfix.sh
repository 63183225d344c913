#!/bin/bash
mkdir -p gpurun_out
{
for cfg in c3 c5; do
  GSP_CHOL_ALGO=recursive python tools/gpu_dist.py 1 $cfg
  for pb in 4 8; do GSP_CHOL_ALGO=panel GSP_CHOL_PB=$pb python tools/gpu_dist.py 1 $cfg; done
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_gemmfix_1gpu.log
GSP_PROF_TIMELINE=1 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python - > gpurun_out/r2_timeline_c3_panel_1gpu.txt 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import gsp_b200 as gsp
lib = gsp.Library()
A = np.zeros((3, 3)); A[0, 0] = A[1, 1] = 1 / 20.0
rng = np.random.default_rng(3)
dinds = np.sort(rng.choice(16384, 1000, replace=False)); z1 = rng.standard_normal(1000) * 0.5
dom = (gsp._lib.make_grid_domain((128, 128), [0.0, 0.0], [1.0, 1.0]), None)
gsp.LUPlan(lib, [(2, 1.0, A)], dom, dinds + 1, z1, 0.0).close()
lib.profile_enable(True)
p = gsp.LUPlan(lib, [(2, 1.0, A)], dom, dinds + 1, z1, 0.0)
print(lib.profile_read(), p.times())
PY
tail -2 gpurun_out/r2_timeline_c3_panel_1gpu.txt | cut -c1-600
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf or lusim and not full" 2>&1 | tail -2
