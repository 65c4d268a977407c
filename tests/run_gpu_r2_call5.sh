#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tests/debug_patch_embed.py > $O/c5_dbg_pe.log 2>&1; echo "exit $?" >> $O/c5_dbg_pe.log
DBG_PDL=1 timeout 120 python tests/debug_patch_embed.py > $O/c5_dbg_pe_pdl.log 2>&1; echo "exit $?" >> $O/c5_dbg_pe_pdl.log
for t in test_single_kernel_patch_embed_matches_im2col_gemm test_fused_residual_and_dsum_switches; do
  timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k $t > $O/c5_variants_$t.log 2>&1
  echo "pytest exit $?" >> $O/c5_variants_$t.log
done
PASST_B200_FUSE_PE=0 python -m pytest tests -m gpu -q -rA --timeout=1500 --deselect tests/test_gpu_variants.py::test_single_kernel_patch_embed_matches_im2col_gemm > $O/c5_pytest_gpu_pe0.log 2>&1
echo "pytest exit $?" >> $O/c5_pytest_gpu_pe0.log
PASST_B200_FUSE_PE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_graphed.py tests/test_gpu_boundary.py -m gpu -q -rA > $O/c5_pytest_pe1.log 2>&1
echo "pytest exit $?" >> $O/c5_pytest_pe1.log
python tests/ncu_kernels.py time > $O/c5_kernels_time_prefetch1.txt 2>&1
PASST_B200_ATTN_PREFETCH=0 python tests/ncu_kernels.py time > $O/c5_kernels_time_prefetch0.txt 2>&1
PASST_B200_FUSE_PE=0 PASST_B200_FUSE_RESID=1 bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/c5_launches_resid1.txt
PASST_B200_FUSE_PE=1 PASST_B200_FUSE_RESID=0 bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/c5_launches_pe1.txt
B="--steps 20 --warmup 5 --stock 0"
for rep in a b; do
  PASST_B200_FUSE_PE=0 PASST_B200_FUSE_RESID=0 timeout 300 python bench.py $B > $O/c5_bench_base_$rep.json 2> $O/c5_bench_base_$rep.err
  PASST_B200_FUSE_PE=0 PASST_B200_FUSE_RESID=1 timeout 300 python bench.py $B > $O/c5_bench_resid1_$rep.json 2> $O/c5_bench_resid1_$rep.err
  PASST_B200_FUSE_PE=1 PASST_B200_FUSE_RESID=0 timeout 300 python bench.py $B > $O/c5_bench_pe1_$rep.json 2> $O/c5_bench_pe1_$rep.err
  PASST_B200_FUSE_PE=0 PASST_B200_FUSE_RESID=0 PASST_B200_ATTN_PREFETCH=0 timeout 300 python bench.py $B > $O/c5_bench_nopf_$rep.json 2> $O/c5_bench_nopf_$rep.err
done
cat $O/c5_dbg_pe.log | tail -8; tail -3 $O/c5_pytest_gpu_pe0.log; tail -3 $O/c5_pytest_pe1.log
cat $O/c5_kernels_time_prefetch1.txt $O/c5_kernels_time_prefetch0.txt | grep attn
for f in $O/c5_bench_*.json; do echo "$f $(head -c 130 $f | cut -c60-130)"; done
