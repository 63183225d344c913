#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu_v1.json 2> gpurun_out/r2_bench_2gpu_v1.err
grep -E "bench\]|Error|error|MiB" gpurun_out/r2_bench_2gpu_v1.err | head -20
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_v1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, d['e2e']['value'], d['e2e']['host_copy_ceiling'], d['e2e']['resident_statistics_variant'])
if 'error' in d['lusim']: print(d['lusim'])
for k in ('c3', 'c5'):
    print(k, {a: d['lusim'][k][a] for a in ('n_devices', 'plan_wall_s', 'factor_device_ms', 'sample_resident_wall_ms', 'realizations_per_s_end_to_end', 'factor_plus_sample_frac_of_peak')}, d['lusim'][k].get('speedup_vs_n1'))
PY
