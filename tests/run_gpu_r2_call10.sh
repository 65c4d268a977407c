#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tests/debug_patch_embed.py > $O/c10_dbg_pe.log 2>&1; echo "exit $?" >> $O/c10_dbg_pe.log
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k test_single_kernel_patch_embed > $O/c10_variants_pe.log 2>&1; echo "pytest exit $?" >> $O/c10_variants_pe.log
PASST_B200_FUSE_PE=1 bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/c10_launches_pe1.txt
tail -7 $O/c10_dbg_pe.log; tail -3 $O/c10_variants_pe.log; grep -E "^one|patch_embed|im2col|gemm2_kernel<2" $O/c10_launches_pe1.txt
