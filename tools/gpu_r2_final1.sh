#!/bin/bash
# final-tree 1-GPU call: full -m gpu suite, default bench, ncu --set full of the new kernels, ncu launch list of a short bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2_pytest_gpu_final.log 2>&1
tail -14 gpurun_out/r2_pytest_gpu_final.log
( time python bench.py ) > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err
tail -c 300 gpurun_out/r2_bench_1gpu_final.err
ncu --set full --clock-control none --import-source on -k regex:potrf_square_kernel -s 4 -c 2 -o gpurun_out/r2_ncu_potrf_square python tools/gpu_lusim.py > gpurun_out/r2_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_dmma_kernel<(\(int\))?1, (\(bool\))?(0|false), (\(int\))?128' -s 6 -c 2 -o gpurun_out/r2_ncu_gemm_update_panel python tools/gpu_lusim.py > gpurun_out/r2_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --reals-per-gpu 16 --e2e-reals 2 --skip-cpu --skip-ensemble > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_bench.csv
