mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r02_pytest_gpu.log | tail -3
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee gpurun_out/r02_smoke.log
( timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1 ) > gpurun_out/r02_bench_default.json
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_default.json').read()); print('default bench', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['config']['cuda_graph'], 'launches', j['gpu_launches'], 'roofline', round(j['roofline']['frac'],3), 'clocks', j['clocks'])"
