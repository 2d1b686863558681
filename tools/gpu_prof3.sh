#!/bin/bash
# kernel-level pass: GPU tests, CUDA-event timings of the hot kernels at bench shapes, ncu --set full captures of them
# (only text exports of the reports come back: the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; tail -15 gpurun_out/pytest_full.log; fi
timeout 300 python tools/prof_kernels.py --time 2>&1 | tee gpurun_out/kernel_times.log
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-gemm_tc|relpos}" -o /tmp/prof_kernels -f python tools/prof_kernels.py > gpurun_out/ncu_kernels.log 2>&1
  echo "ncu rc=$?"; tail -2 gpurun_out/ncu_kernels.log
  ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
  ncu -i /tmp/prof_kernels.ncu-rep --page source --csv -k regex:"relpos_attn_bwd_q_tc" > gpurun_out/prof_src_bwd_q.csv 2>/dev/null
  ncu -i /tmp/prof_kernels.ncu-rep --page source --csv -k regex:"relpos_attn_fwd_tc" > gpurun_out/prof_src_fwd.csv 2>/dev/null
  ncu -i /tmp/prof_kernels.ncu-rep --page source --csv -k regex:"relpos_attn_bwd_pos" > gpurun_out/prof_src_bwd_pos.csv 2>/dev/null
  ncu -i /tmp/prof_kernels.ncu-rep --page source --csv -k regex:"gemm_tc" -c 1 > gpurun_out/prof_src_gemm_silu.csv 2>/dev/null
  gzip -f gpurun_out/prof_src_*.csv; ls -la gpurun_out
fi
