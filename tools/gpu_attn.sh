#!/bin/bash
# Attention kernels first (A/B against the legacy mma.sync key-side kernel when something fails), then the full round.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k relpos_attention > gpurun_out/attn_tc.log 2>&1
tail -15 gpurun_out/attn_tc.log
if ! grep -q " passed" gpurun_out/attn_tc.log || grep -q "failed" gpurun_out/attn_tc.log; then
  B2C_ATTN_IMPL=mmakv timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k relpos_attention > gpurun_out/attn_mmakv.log 2>&1
  echo "--- mmakv"; tail -15 gpurun_out/attn_mmakv.log
fi
bash tools/gpu_round.sh
