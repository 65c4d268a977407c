#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
NCU_ONLY=patch_embed timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f \
   -k regex:patch_embed -o $O/c15_pe python tests/ncu_kernels.py > $O/c15_ncu.log 2>&1
tail -3 $O/c15_ncu.log
