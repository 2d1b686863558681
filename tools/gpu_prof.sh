#!/bin/bash
# ncu --set full captures of the dominant kernels (one bench step each) + a focused test log
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -s -k "fastconformer or golden" > gpurun_out/test_fastconf.log 2>&1
echo "fastconf rc=$?"
grep -E "FAIL|Error|passed|failed" gpurun_out/test_fastconf.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 150 -c 24 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:relpos_attn -s 60 -c 8 -o gpurun_out/prof_attn -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out
