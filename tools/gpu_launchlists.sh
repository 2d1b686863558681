#!/bin/bash
mkdir -p gpurun_out
for c in 2 3; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/launches_cfg$c.csv \
    python bench.py --config $c --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_cfg$c.log 2>&1
  echo "ncu cfg$c rc=$?"
done
