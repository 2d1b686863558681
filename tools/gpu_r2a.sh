#!/bin/bash
# round 2, first GPU pass: new full-depth / C=1024 / dither tests first (they are the unknowns), then the old suite,
# then bench lines for every BASELINE config.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -s > gpurun_out/r2a_fulldepth.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error" gpurun_out/r2a_fulldepth.log | tail -60
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -s > gpurun_out/r2a_pytest_rest.log 2>&1
grep -E "FAIL|passed|failed|Error" gpurun_out/r2a_pytest_rest.log | tail -15
for c in 2 3 4 5; do
  ( timeout 600 python bench.py --config $c --steps 5 --warmup 3 2>&1 | tail -3 ) > gpurun_out/r2a_bench_cfg$c.log
  tail -c 1500 gpurun_out/r2a_bench_cfg$c.log; echo
done
