#!/bin/bash
# round-2 GPU call 1: full GPU test suite (incl. the new full-depth gradient parity), ncu --set full of the non-GEMM
# kernels, standalone timings.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --timeout=1500 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
python tests/ncu_kernels.py time > gpurun_out/r2_kernels_time.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
    -o gpurun_out/r2_kernels python tests/ncu_kernels.py > gpurun_out/r2_ncu_kernels.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
cat gpurun_out/r2_kernels_time.txt
