#!/bin/bash
# round 2, third GPU pass: parity mode + CTC + graph tests, ncu launch lists of one eager step for configs 2/3/4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_runtime_gpu.py -m gpu -q -s > gpurun_out/r2c_runtime.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error|captured" gpurun_out/r2c_runtime.log | tail -30
timeout 1500 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -s -k "parity or dither" > gpurun_out/r2c_parity.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error|dither" gpurun_out/r2c_parity.log | tail -60
for c in 2 3 4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches_cfg$c.csv \
    python bench.py --config $c --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r2c_ncu_cfg$c.log 2>&1
  echo "ncu cfg$c rc=$?"
  python tools/launch_table.py gpurun_out/launches_cfg$c.csv "r02 cfg$c launch list" | head -40
done
