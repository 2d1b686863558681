#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<256, 0, 0, [12]" -s 40 -c 4 -o gpurun_out/prof_gemm2 -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu gemm rc=$?"; tail -3 gpurun_out/ncu_gemm2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd|dwconv_bwd" -s 20 -c 2 -o gpurun_out/prof_ln -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ln.log 2>&1
echo "ncu ln rc=$?"
ls -la gpurun_out | head -20
