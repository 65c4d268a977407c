#!/bin/bash
# round-2 GPU call 2: variant tests (one process each), fixed tests on the round-1 schedule, full suite with the new
# defaults (single-kernel patch embedding tested separately), launch lists of one step for both schedules, bench A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
R1="PASST_B200_PDL=0 PASST_B200_ATTN_FWD=1 PASST_B200_FUSE_RESID=0 PASST_B200_FUSE_DSUM=0 PASST_B200_FUSE_PE=0"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/c2_smi.txt 2>&1
for t in test_single_kernel_patch_embed_matches_im2col_gemm test_pdl_on_off_same_results test_fused_residual_and_dsum_switches test_attention_forward_variants_agree; do
  PASST_B200_FUSE_PE=0 timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k $t > $O/c2_variants_$t.log 2>&1
  echo "pytest exit $?" >> $O/c2_variants_$t.log
done
env $R1 timeout 900 python -m pytest tests/test_gpu_fulldepth.py tests/test_gpu_parity.py -m gpu -q -rA > $O/c2_pytest_r1sched_fixed.log 2>&1
echo "pytest exit $?" >> $O/c2_pytest_r1sched_fixed.log
mkdir -p $O/parity_r1sched && cp $O/parity_grads_*.txt $O/parity_r1sched/ 2>/dev/null
PASST_B200_FUSE_PE=0 python -m pytest tests -m gpu -q -rA --timeout=1500 > $O/c2_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/c2_pytest_gpu.log
PASST_B200_FUSE_PE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graphed.py -m gpu -q -rA > $O/c2_pytest_pe1.log 2>&1
echo "pytest exit $?" >> $O/c2_pytest_pe1.log
# launch lists of one train step (ncu, serialised): round-1 schedule vs new defaults
env $R1 bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/c2_launches_r1sched.txt
PASST_B200_FUSE_PE=0 bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/c2_launches_new.txt
B="--steps 20 --warmup 5"
PASST_B200_FUSE_PE=0 timeout 900 python bench.py $B > $O/c2_bench_cfg2.json 2> $O/c2_bench_cfg2.err
env $R1 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_r1sched.json 2> $O/c2_bench_cfg2_r1sched.err
PASST_B200_FUSE_PE=0 PASST_B200_PDL=0 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_nopdl.json 2> $O/c2_bench_cfg2_nopdl.err
PASST_B200_FUSE_PE=0 PASST_B200_ATTN_FWD=1 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_attn1.json 2> $O/c2_bench_cfg2_attn1.err
PASST_B200_FUSE_PE=0 PASST_B200_FUSE_RESID=0 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_noresid.json 2> $O/c2_bench_cfg2_noresid.err
PASST_B200_FUSE_PE=0 PASST_B200_FUSE_DSUM=0 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_nodsum.json 2> $O/c2_bench_cfg2_nodsum.err
PASST_B200_FUSE_PE=1 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_pe1.json 2> $O/c2_bench_cfg2_pe1.err
env $R1 timeout 300 python bench.py $B --stock 0 > $O/c2_bench_cfg2_r1sched_b.json 2> $O/c2_bench_cfg2_r1sched_b.err
for c in cfg1 cfg3 cfg4 cfg5; do
  PASST_B200_FUSE_PE=0 timeout 600 python bench.py --config $c $B --stock 0 > $O/c2_bench_$c.json 2> $O/c2_bench_$c.err
done
PASST_B200_FUSE_PE=0 timeout 300 python bench.py --config cfg1 --precision fp32 --steps 10 --warmup 3 --stock 0 > $O/c2_bench_cfg1_fp32.json 2> $O/c2_bench_cfg1_fp32.err
PASST_B200_FUSE_PE=0 timeout 600 python bench.py --config cfg4 $B --stock 1 > $O/c2_bench_cfg4_stock.json 2> $O/c2_bench_cfg4_stock.err
tail -3 $O/c2_pytest_gpu.log
for f in $O/c2_bench_*.json; do echo "$f $(head -c 160 $f)"; done
