#!/bin/bash
# ncu --set full capture of the hot kernels at the bench shapes (text export only: the .ncu-rep exceeds the return limit)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|relpos|layernorm_bwd|dwconv|mel_frames|mel_normalize" -o /tmp/prof_kernels -f python tools/prof_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_kernels.log
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/prof_kernels_raw.csv
