#!/bin/bash
# round-2 GPU call 1: full GPU test suite (incl. the new full-depth gradient parity), ncu --set full of the non-GEMM
# kernels, standalone timings, bench lines of every BASELINE.json configuration.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
python -m pytest tests -m gpu -q -rA --timeout=1500 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
python tests/ncu_kernels.py time > gpurun_out/r2_kernels_time.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
    -o gpurun_out/r2_kernels python tests/ncu_kernels.py > gpurun_out/r2_ncu_kernels.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --stock 0 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
timeout 300 python bench.py --config cfg1 --precision fp32 --steps 10 --warmup 3 --stock 0 > gpurun_out/r2_bench_cfg1_fp32.json 2> gpurun_out/r2_bench_cfg1_fp32.err
tail -5 gpurun_out/r2_pytest_gpu.log
cat gpurun_out/r2_kernels_time.txt
head -c 600 gpurun_out/r2_bench_cfg2.json
