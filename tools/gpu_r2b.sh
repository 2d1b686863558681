#!/bin/bash
# round 2, second GPU pass: runtime tests (prep_weights, fused AdamW, CUDA graph), full suite, bench all configs graph vs eager
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_runtime_gpu.py -m gpu -q -s -x > gpurun_out/r2b_runtime.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error|captured" gpurun_out/r2b_runtime.log | tail -30
timeout 1500 python -m pytest tests/test_full_depth_gpu.py tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q > gpurun_out/r2b_pytest_rest.log 2>&1
grep -E "FAIL|passed|failed|Error" gpurun_out/r2b_pytest_rest.log | tail -15
for c in 2 3 4; do
  ( timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/r2b_bench_cfg$c.log
  ( timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-graph 2>&1 | tail -3 ) > gpurun_out/r2b_bench_cfg${c}_eager.log
  python - <<PY
import json
for f in ("gpurun_out/r2b_bench_cfg$c.log","gpurun_out/r2b_bench_cfg${c}_eager.log"):
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; j=json.loads(l)
        print(f, 'value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']),'graph',j['config']['cuda_graph'],'launches',j['gpu_launches'],'gemm frac',round(j['roofline']['frac'],3), 'step_ms', j['config']['step_ms'])
    except Exception as e:
        print(f,'ERR',e); print(open(f).read()[-1500:])
PY
done
