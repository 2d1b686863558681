#!/bin/bash
# round 2, fourth GPU pass: full suite after the mel / depthwise / optimizer changes, bench all configs (graph)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1
grep -E "FAIL|passed|failed|Error|error" gpurun_out/r2d_pytest.log | tail -15
grep -E "^\[mel" gpurun_out/r2d_pytest.log | head
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k mel 2>&1 | grep -E "^\[mel|passed|failed"
for c in 2 3 4 5; do
  ( timeout 900 python bench.py --config $c --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/r2d_bench_cfg$c.log
  python - <<PY
import json
f="gpurun_out/r2d_bench_cfg$c.log"
try:
    l=[x for x in open(f) if x.startswith('{')][-1]; j=json.loads(l)
    print(f, 'value',round(j['value']),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value']),'launches',j['gpu_launches'],'roofline frac',round(j['roofline']['frac'],3), j['config'].get('cuda_graph'), 'eager', (j.get('gpu_eager_baseline') or {}).get('value'), 'cpu', (j.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print(f,'ERR',e); print(open(f).read()[-1500:])
PY
done
