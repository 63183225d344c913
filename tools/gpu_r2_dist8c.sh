#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multigpu_8c.log 2>&1
tail -4 gpurun_out/r2_pytest_multigpu_8c.log
{
for cfg in c5 c3; do
  for pb in 4 8; do GSP_CHOL_ALGO=panel GSP_CHOL_PB=$pb python tools/gpu_dist.py 8 $cfg 3; done
  echo "unfused:"; GSP_CHOL_FUSED_SQUARE=0 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 8 $cfg 3
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 4 $cfg 3
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_dist_sweep_8gpu_c.log
GSP_PROF_TIMELINE=1 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python - > gpurun_out/r2_timeline_c5_8gpu_c.txt 2>&1 <<'PY'
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
tail -1 gpurun_out/r2_timeline_c5_8gpu_c.txt | cut -c1-500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu_v1.json 2> gpurun_out/r2_bench_8gpu_v1.err
grep -E "bench\]|Error|error" gpurun_out/r2_bench_8gpu_v1.err | head -12
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_8gpu_v1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, d['e2e']['value'], d['e2e']['host_copy_ceiling']['pinned_h2d_plus_d2h_GBps_all_ranks'], d['e2e']['resident_statistics_variant']['value'])
if 'error' in d['lusim']: print(d['lusim'])
for k in ('c3', 'c5'):
    print(k, {a: d['lusim'][k][a] for a in ('n_devices', 'plan_wall_s', 'factor_device_ms', 'sample_resident_wall_ms', 'realizations_per_s_end_to_end', 'realizations_per_s_plan_plus_resident_sampling', 'factor_plus_sample_frac_of_peak')}, d['lusim'][k].get('speedup_vs_n1'))
PY
