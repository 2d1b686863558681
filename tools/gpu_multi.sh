#!/bin/bash
# multi-GPU pass: bash tools/gpu_multi.sh N   (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, bench args...
  name=$1; shift
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N "$@" 2>&1 | grep '^{' | tail -1 ) > gpurun_out/multi_n${N}_$name.json
  python - <<PY
import json
f="gpurun_out/multi_n${N}_$name.json"
try:
    j=json.loads(open(f).read())
    print("N=$N $name:", 'value', round(j['value']), 'ms', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['config'].get('cuda_graph'), 'scaling', j['scaling'], 'global_batch', j['config']['global_batch'], 'gemm frac', round(j['roofline']['frac'],3))
except Exception as e:
    print("N=$N $name: ERR", e)
PY
}
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s 2>&1 | tail -4
  run cfg2_weak --config 2 --steps 10 --warmup 3
  run cfg2_weak_eager --config 2 --steps 10 --warmup 3 --no-graph
  run cfg2_strong --config 2 --scaling strong --steps 10 --warmup 3
fi
if [ "$N" = "4" ]; then
  run cfg3_weak --config 3 --steps 10 --warmup 3
  run cfg2_strong --config 2 --scaling strong --steps 10 --warmup 3
fi
if [ "$N" = "8" ]; then
  run cfg4_weak --config 4 --steps 10 --warmup 3
  run cfg2_strong --config 2 --scaling strong --steps 10 --warmup 3
fi
