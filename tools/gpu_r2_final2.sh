#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err
grep -E "Error|error" gpurun_out/r2_bench_2gpu_final.err | head -5
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step')}, d['e2e']['value'], d.get('fftsim_c2', {}).get('realizations_per_s'))
if 'error' in d['lusim']: print(d['lusim'])
for k in ('c3', 'c5'):
    print(k, {a: d['lusim'][k][a] for a in ('n_devices', 'plan_wall_s', 'factor_device_ms', 'sample_resident_wall_ms', 'realizations_per_s_end_to_end')}, d['lusim'][k]['roofline']['frac'], d['lusim'][k].get('speedup_vs_n1'))
PY
python bench.py --impl reference --steps 2 --warmup 1 --skip-lusim | cut -c1-400
