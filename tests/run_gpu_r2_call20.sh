#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -x -k test_attention_backward_variants > $O/c20_bwd_variants.log 2>&1; echo "pytest exit $?" >> $O/c20_bwd_variants.log
NCU_ONLY=attn_bwd timeout 120 python tests/ncu_kernels.py time > $O/c20_bwd_time_v1.txt 2>&1
PASST_B200_ATTN_BWD=2 NCU_ONLY=attn_bwd timeout 120 python tests/ncu_kernels.py time > $O/c20_bwd_time_v2.txt 2>&1
tail -12 $O/c20_bwd_variants.log; cat $O/c20_bwd_time_v1.txt $O/c20_bwd_time_v2.txt
