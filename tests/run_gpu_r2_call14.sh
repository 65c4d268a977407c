#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tests/debug_patch_embed.py > $O/c14_dbg_pe.log 2>&1; echo "exit $?" >> $O/c14_dbg_pe.log
DBG_PDL=1 timeout 120 python tests/debug_patch_embed.py > $O/c14_dbg_pe_pdl.log 2>&1; echo "exit $?" >> $O/c14_dbg_pe_pdl.log
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k test_single_kernel_patch_embed > $O/c14_variants_pe.log 2>&1; echo "pytest exit $?" >> $O/c14_variants_pe.log
NCU_ONLY=im2col,patch_embed,pe_gemm python tests/ncu_kernels.py time > $O/c14_pe_time.txt 2>&1
tail -4 $O/c14_dbg_pe.log; tail -2 $O/c14_dbg_pe_pdl.log; tail -3 $O/c14_variants_pe.log; cat $O/c14_pe_time.txt
