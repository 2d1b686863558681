#!/bin/bash
# A/B at 2 GPUs: how many SMs NCCL may take for the gradient all-reduce that overlaps the backward (NCCL_MAX_CTAS)
mkdir -p gpurun_out
for v in default 2 4 8; do
  if [ "$v" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --config 2 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 ) > gpurun_out/nccl_ctas_$v.json
  python -c "
import json; j=json.loads(open('gpurun_out/nccl_ctas_$v.json').read()); print('NCCL_MAX_CTAS=$v', round(j['value']), round(j['ms_per_step'],3), 'gemm frac', round(j['roofline']['frac'],3), j['config']['step_ms'])"
done
