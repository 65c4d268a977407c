#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU_ONLY=im2col,patch_embed,pe_gemm python tests/ncu_kernels.py time > $O/c13_pe_time.txt 2>&1
NCU_ONLY=patch_embed timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
   -k regex:patch_embed -o $O/c13_pe python tests/ncu_kernels.py > $O/c13_ncu.log 2>&1
cat $O/c13_pe_time.txt; tail -3 $O/c13_ncu.log
