#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -x -k test_attention_forward_variants > $O/c23_fwd_variants.log 2>&1; echo "pytest exit $?" >> $O/c23_fwd_variants.log
for v in 2 4; do PASST_B200_ATTN_FWD=$v NCU_ONLY=attn_fwd timeout 120 python tests/ncu_kernels.py time 2>&1 | grep attn_fwd > $O/c23_fwd_time_v$v.txt; done
tail -12 $O/c23_fwd_variants.log; cat $O/c23_fwd_time_v2.txt $O/c23_fwd_time_v4.txt
