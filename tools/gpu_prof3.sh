#!/bin/bash
# kernel-level pass: GPU tests, CUDA-event timings of the hot kernels at bench shapes, one ncu --set full capture of them
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log
timeout 300 python tools/prof_kernels.py --time 2>&1 | tee gpurun_out/kernel_times.log
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels -f python tools/prof_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_kernels.log; ls -la gpurun_out/prof_kernels.ncu-rep
