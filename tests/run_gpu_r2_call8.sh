#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
PASST_B200_ATTN_POLYEXP=1 timeout 300 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k test_attention_forward_variants_agree > $O/c8_variants_poly.log 2>&1; echo "pytest exit $?" >> $O/c8_variants_poly.log
for pe in 0 1; do for pf in 0 1; do
  PASST_B200_ATTN_POLYEXP=$pe PASST_B200_ATTN_PREFETCH=$pf python tests/ncu_kernels.py time 2>&1 | grep attn_fwd | sed "s/^/poly=$pe prefetch=$pf /" >> $O/c8_attn_time.txt
done; done
NCU_BATCH=256 PASST_B200_ATTN_POLYEXP=0 python tests/ncu_kernels.py time 2>&1 | grep attn_fwd | sed "s/^/B=256 poly=0 /" >> $O/c8_attn_time.txt
NCU_BATCH=256 PASST_B200_ATTN_POLYEXP=1 python tests/ncu_kernels.py time 2>&1 | grep attn_fwd | sed "s/^/B=256 poly=1 /" >> $O/c8_attn_time.txt
PASST_B200_ATTN_POLYEXP=1 timeout 300 ncu --set full --clock-control none --profile-from-start off -f -k regex:attn_fwd2 -o $O/c8_attn2_poly python tests/ncu_kernels.py > $O/c8_ncu.log 2>&1
tail -3 $O/c8_variants_poly.log; cat $O/c8_attn_time.txt
