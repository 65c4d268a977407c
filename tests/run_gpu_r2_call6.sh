#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -rA -k test_attention_forward_variants_agree > $O/c6_variants_attn.log 2>&1; echo "pytest exit $?" >> $O/c6_variants_attn.log
for v in 1 2 3; do PASST_B200_ATTN_FWD=$v python tests/ncu_kernels.py time 2>&1 | grep attn_fwd > $O/c6_attn_time_v$v.txt; done
PASST_B200_ATTN_FWD=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fulldepth.py -m gpu -q -rA > $O/c6_pytest_attn3.log 2>&1; echo "pytest exit $?" >> $O/c6_pytest_attn3.log
PASST_B200_ATTN_FWD=3 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k regex:attn_fwd3 -o $O/c6_attn3 python tests/ncu_kernels.py > $O/c6_ncu.log 2>&1
B="--steps 20 --warmup 5 --stock 0"
for rep in a b; do
  PASST_B200_ATTN_FWD=2 timeout 300 python bench.py $B > $O/c6_bench_attn2_$rep.json 2> $O/c6_bench_attn2_$rep.err
  PASST_B200_ATTN_FWD=3 timeout 300 python bench.py $B > $O/c6_bench_attn3_$rep.json 2> $O/c6_bench_attn3_$rep.err
done
PASST_B200_ATTN_FWD=3 timeout 300 python bench.py --config cfg4 $B > $O/c6_bench_cfg4_attn3.json 2> $O/c6_bench_cfg4_attn3.err
PASST_B200_ATTN_FWD=2 timeout 300 python bench.py --config cfg4 $B > $O/c6_bench_cfg4_attn2.json 2> $O/c6_bench_cfg4_attn2.err
tail -4 $O/c6_variants_attn.log; cat $O/c6_attn_time_v*.txt; tail -3 $O/c6_pytest_attn3.log
for f in $O/c6_bench_*.json; do echo "$f $(head -c 130 $f | cut -c60-130)"; done
