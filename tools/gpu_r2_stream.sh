#!/bin/bash
for v in 1 0; do
GSP_FFT_STREAM=$v python bench.py --steps 4 --warmup 3 --skip-lusim --skip-cpu --skip-ensemble 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stream=$v', round(d['value'],1), d['e2e']['matches_device_path'], d['invariants_ok'], round(d['roofline']['frac'],3))"
done
