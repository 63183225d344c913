#!/bin/bash
# round-2 first GPU call (1 GPU): full -m gpu suite incl. the new full-size oracle tests, default bench, ncu --set full of the
# contraction kernels, compute-sanitizer on the smoke shapes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/r2_gpu.txt
nproc | tee -a gpurun_out/r2_gpu.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -30 gpurun_out/r2_pytest_gpu.log
( time python bench.py ) > gpurun_out/r2_bench_v0.json 2> gpurun_out/r2_bench_v0.err
tail -c 600 gpurun_out/r2_bench_v0.err
# ncu full: update GEMM (128x128 tiles), TRSM GEMM, sample GEMM, diagonal block
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_dmma_kernel<(\(int\))?1, (\(bool\))?(0|false), (\(int\))?128' -s 40 -c 2 -o gpurun_out/r2_ncu_gemm_update python tools/gpu_lusim.py > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_dmma_kernel<(\(int\))?2,' -s 1 -c 1 -o gpurun_out/r2_ncu_gemm_sample python tools/gpu_lusim.py > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu1.log gpurun_out/r2_ncu2.log
ls -la gpurun_out/*.ncu-rep
# compute-sanitizer on the smoke shapes
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck.log
tail -5 gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
