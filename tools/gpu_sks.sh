#!/bin/bash
# A/B: positional-band skew through shared memory (B2C_ATTN_SKEW=smem) vs the register barrel shift in the tcgen05 attention forward
mkdir -p gpurun_out
B2C_ATTN_SKEW=smem timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -k "attention or encoder_forward_backward or golden or end_to_end" 2>&1 | tail -3
echo "== register skew"; timeout 300 python tools/prof_kernels.py --time 2>&1 | grep relpos_attn_fwd
echo "== smem skew"; B2C_ATTN_SKEW=smem timeout 300 python tools/prof_kernels.py --time 2>&1 | grep relpos_attn_fwd
for v in reg smem; do
  ( B2C_ATTN_SKEW=$v timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | grep '^{' | tail -1 ) > gpurun_out/sks_${v}_cfg2.json
  python -c "
import json; j=json.loads(open('gpurun_out/sks_${v}_cfg2.json').read()); print('skew=$v cfg2', round(j['value']), round(j['ms_per_step'],3))"
done
