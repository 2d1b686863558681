#!/bin/bash
# Full GPU check: pytest -m gpu, smoke(), bench (ours), ncu launch list of one bench step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_full.log 2>&1; grep -E "FAIL|passed|failed|Error" gpurun_out/pytest_full.log | tail -25 | tee gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/smoke.log
( timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 2>&1 | tail -5 ) | tee gpurun_out/bench.log
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?" | tee -a gpurun_out/bench.log
fi
