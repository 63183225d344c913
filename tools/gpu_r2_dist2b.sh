#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_multigpu_2b.log 2>&1
tail -4 gpurun_out/r2_pytest_multigpu_2b.log
( GSP_CHOL_ALGO=panel python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "potrf or lusim and not c5_full" ) 2>&1 | tail -2
{
for cfg in c5 c3; do
  for g in 2 1; do
    GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py $g $cfg 3
    GSP_CHOL_FUSED_SQUARE=0 GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py $g $cfg 3
  done
  GSP_CHOL_ALGO=panel GSP_CHOL_PB=8 python tools/gpu_dist.py 2 $cfg 3
done
# repeat the configuration that raced before (G = 2, c5) a few times
for k in 1 2 3; do GSP_CHOL_ALGO=panel GSP_CHOL_PB=4 python tools/gpu_dist.py 2 c5 2; done
} 2>&1 | grep -v "^$" | tee gpurun_out/r2_dist_sweep_2gpu_b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu_v1.json 2> gpurun_out/r2_bench_2gpu_v1.err
tail -c 1500 gpurun_out/r2_bench_2gpu_v1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_v1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, d['e2e']['value'], d['e2e']['host_copy_ceiling'])
for k in ('c3', 'c5'):
    print(k, {a: d['lusim'][k][a] for a in ('n_devices', 'plan_wall_s', 'factor_device_ms', 'sample_resident_wall_ms', 'realizations_per_s_end_to_end', 'factor_plus_sample_frac_of_peak')}, d['lusim'][k].get('speedup_vs_n1'))
PY
