#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shared_factor or lusim_bivariate or rand_api" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err
tail -c 200 gpurun_out/r2_bench_1gpu_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_1gpu_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'])
for k in ('c3', 'c5'):
    c = d['lusim'][k]
    print(k, {a: c.get(a) for a in ('plan_wall_s', 'plan_wall_s_unshared', 'factor_device_ms', 'factor_device_ms_unshared', 'sample_resident_wall_ms', 'realizations_per_s_end_to_end', 'factor_tflops')}, c['roofline']['frac'])
PY
