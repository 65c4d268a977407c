#!/bin/bash
# usage: tests/run_bringup.sh group1 group2 ...   (each group under its own timeout + process)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "=== $g" | tee -a gpurun_out/bringup.log
  timeout 300 python tests/bringup_gpu.py "$g" >> gpurun_out/bringup.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/bringup.log
done
tail -c 6000 gpurun_out/bringup.log
