#!/bin/bash
# final single-GPU validation of the round: full GPU suite, smoke, bench lines of every config, reference arm, kernel timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
grep -E "FAIL|passed|failed|Error" gpurun_out/r02_pytest_gpu.log | tail -12
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/r02_smoke.log
for c in 2 3 4 5; do
  ( timeout 420 python bench.py --config $c --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 ) > gpurun_out/r02_bench_cfg$c.json
  python - <<PY
import json
f="gpurun_out/r02_bench_cfg$c.json"
try:
    j=json.loads(open(f).read())
    print(f, 'value',round(j['value']),'ms',round(j['ms_per_step'],3),'e2e',round(j['e2e']['value']),'launches',j['gpu_launches'],'roofline',round(j['roofline']['frac'],3), j['config'].get('cuda_graph'), 'eager', (j.get('gpu_eager_baseline') or {}).get('value'), 'cpu', (j.get('cpu_baseline') or {}).get('value'), 'traffic', j['roofline'].get('traffic'))
except Exception as e:
    print(f,'ERR',e)
PY
done
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 ) > gpurun_out/r02_bench_reference.json; head -c 600 gpurun_out/r02_bench_reference.json; echo
timeout 300 python tools/prof_kernels.py --time 2>&1 | tee gpurun_out/kernel_times.log | tail -22
