#!/bin/bash
# Runs each GPU test group in its own process under a timeout so one hang / crash does not hide the others.
# Usage (under gpurun): bash tools/gpu_diag.sh [pytest -k expressions...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then
  groups=("gemm_kmajor" "gemm_mn_major" "gemm_epilogue or gemm_dropout" "mel" "layernorm or colsum or relpos_table" "conv_module" "attention")
fi
i=0
for g in "${groups[@]}"; do
  i=$((i+1))
  echo "=== group $i: $g" | tee -a gpurun_out/diag_summary.txt
  timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -s -k "$g" > "gpurun_out/diag_$i.log" 2>&1
  rc=$?
  echo "rc=$rc" | tee -a gpurun_out/diag_summary.txt
  grep -E "^\[|passed|failed|error|Error|timeout|rel err|assert" "gpurun_out/diag_$i.log" | tail -60 >> gpurun_out/diag_summary.txt
done
tail -150 gpurun_out/diag_summary.txt
