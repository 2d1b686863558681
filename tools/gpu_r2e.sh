#!/bin/bash
# round 2, fifth GPU pass: parity-mode backward, depthwise fix, kernel timings + ncu --set full capture of the hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_runtime_gpu.py -m gpu -q -s -k "seed_counter" > gpurun_out/r2e_runtime.log 2>&1
grep -E "passed|failed|captured" gpurun_out/r2e_runtime.log | tail -5
timeout 1500 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -s -k "gradients or train_statistics" > gpurun_out/r2e_parity_grad.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error" gpurun_out/r2e_parity_grad.log | tail -30
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py -m gpu -q -k "subsampl or fastconformer or dw" > gpurun_out/r2e_dw.log 2>&1
tail -3 gpurun_out/r2e_dw.log
( timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2e_bench_cfg3.log
python -c "
import json
j=json.loads([x for x in open('gpurun_out/r2e_bench_cfg3.log') if x.startswith('{')][-1]); print('cfg3', round(j['value']), round(j['ms_per_step'],3))"
timeout 600 python tools/prof_kernels.py --time 2>&1 | tee gpurun_out/kernel_times.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|relpos|layernorm_bwd|dwconv_bwd|dwconv_fwd|mel_frames|mel_normalize" -o /tmp/prof_kernels -f python tools/prof_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_kernels.log
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/prof_kernels_raw.csv
