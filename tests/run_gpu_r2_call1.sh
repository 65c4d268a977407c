#!/bin/bash
# round-2 GPU call: (1) full GPU suite on the round-1 kernel schedule (all new switches off), (2) the A/B variant tests,
# one process each, (3) full suite with the new defaults, (4) ncu --set full of the non-GEMM kernels, (5) bench lines:
# cfg2 with the stock arms, cfg2 with each switch turned off, the other BASELINE.json configurations.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
PASST_B200_PDL=0 PASST_B200_ATTN_FWD=1 PASST_B200_FUSE_RESID=0 PASST_B200_FUSE_DSUM=0 PASST_B200_FUSE_PE=0 \
  python -m pytest tests -m gpu -q -rA --timeout=1500 --deselect tests/test_gpu_variants.py > $O/r2_pytest_gpu_base.log 2>&1
echo "pytest exit $?" >> $O/r2_pytest_gpu_base.log
for t in test_attention_forward_variants_agree test_pdl_on_off_same_results test_fused_residual_and_dsum_switches test_single_kernel_patch_embed_matches_im2col_gemm; do
  timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k $t > $O/r2_pytest_variants_$t.log 2>&1
  echo "pytest exit $?" >> $O/r2_pytest_variants_$t.log
done
python -m pytest tests -m gpu -q -rA --timeout=1500 > $O/r2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/r2_pytest_gpu.log
python tests/ncu_kernels.py time > $O/r2_kernels_time.txt 2>&1
PASST_B200_ATTN_FWD=1 python tests/ncu_kernels.py time > $O/r2_kernels_time_attn1.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
    -o $O/r2_kernels python tests/ncu_kernels.py > $O/r2_ncu_kernels.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_cfg2.json 2> $O/r2_bench_cfg2.err
PASST_B200_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_nopdl.json 2> $O/r2_bench_cfg2_nopdl.err
PASST_B200_ATTN_FWD=1 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_attn1.json 2> $O/r2_bench_cfg2_attn1.err
PASST_B200_FUSE_RESID=0 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_noresid.json 2> $O/r2_bench_cfg2_noresid.err
PASST_B200_FUSE_DSUM=0 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_nodsum.json 2> $O/r2_bench_cfg2_nodsum.err
PASST_B200_FUSE_PE=0 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_nope.json 2> $O/r2_bench_cfg2_nope.err
PASST_B200_PDL=0 PASST_B200_ATTN_FWD=1 PASST_B200_FUSE_RESID=0 PASST_B200_FUSE_DSUM=0 PASST_B200_FUSE_PE=0 timeout 300 python bench.py --steps 20 --warmup 5 --stock 0 > $O/r2_bench_cfg2_r1sched.json 2> $O/r2_bench_cfg2_r1sched.err
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --stock 0 > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
done
timeout 300 python bench.py --config cfg1 --precision fp32 --steps 10 --warmup 3 --stock 0 > $O/r2_bench_cfg1_fp32.json 2> $O/r2_bench_cfg1_fp32.err
tail -3 $O/r2_pytest_gpu_base.log; tail -3 $O/r2_pytest_gpu.log
cat $O/r2_kernels_time.txt
head -c 400 $O/r2_bench_cfg2.json
