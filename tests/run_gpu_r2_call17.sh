#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tests/debug_patch_embed.py > $O/c17_dbg_pe.log 2>&1; echo "exit $?" >> $O/c17_dbg_pe.log
NCU_ONLY=im2col,patch_embed,pe_gemm python tests/ncu_kernels.py time > $O/c17_pe_time.txt 2>&1
NCU_ONLY=patch_embed timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
   -k regex:patch_embed -o $O/c17_pe python tests/ncu_kernels.py > $O/c17_ncu.log 2>&1
tail -3 $O/c17_dbg_pe.log; cat $O/c17_pe_time.txt; tail -2 $O/c17_ncu.log
