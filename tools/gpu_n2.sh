mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_full_depth_gpu.py -m gpu -q -s -k "gradients" > gpurun_out/r2f_parity_grad.log 2>&1
grep -E "^\[|FAIL|passed|failed|Error|error" gpurun_out/r2f_parity_grad.log | tail -30
bash tools/gpu_multi.sh 2
