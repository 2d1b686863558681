#!/bin/bash
# A/B: 4 softmax threads per query row (16 softmax warps per CTA) instead of 2 in the tcgen05 attention kernels
mkdir -p gpurun_out
echo "== NT=2 (default)"; timeout 300 python tools/prof_kernels.py --time 2>&1 | grep relpos
echo "== NT=4"; B2C_ATTN_NT=4 timeout 300 python tools/prof_kernels.py --time 2>&1 | grep relpos
B2C_ATTN_NT=4 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -k "attention or encoder_forward_backward or golden" 2>&1 | tail -3
for nt in 2 4; do
  ( B2C_ATTN_NT=$nt timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | grep '^{' | tail -1 ) > gpurun_out/nt${nt}_cfg2.json
  python -c "
import json; j=json.loads(open('gpurun_out/nt${nt}_cfg2.json').read()); print('NT=$nt cfg2', round(j['value']), round(j['ms_per_step'],3))"
  ( B2C_ATTN_NT=$nt timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | grep '^{' | tail -1 ) > gpurun_out/nt${nt}_cfg3.json
  python -c "
import json; j=json.loads(open('gpurun_out/nt${nt}_cfg3.json').read()); print('NT=$nt cfg3', round(j['value']), round(j['ms_per_step'],3))"
done
