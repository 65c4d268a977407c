#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="--steps 20 --warmup 5 --stock 0"
run8() { tag=$1; port=$2; shift 2; timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 $B > $O/s8b_$tag.json 2> $O/s8b_$tag.err; }
run8 r0_default 29711 PASST_DDP_RESERVE=0 NCCL_MAX_CTAS=32
run8 r8 29712 PASST_DDP_RESERVE=8 NCCL_MAX_CTAS=8
run8 r16 29713 PASST_DDP_RESERVE=16 NCCL_MAX_CTAS=16
run8 r4_dbg 29714 PASST_DDP_RESERVE=4 NCCL_MAX_CTAS=4 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL
grep -E "NCCL INFO.*(Channel|channels|NVLS|Algo|Trees|Ring|nChannels|comm 0x.* rank 0)" $O/s8b_r4_dbg.err | head -30 > $O/s8b_nccl_info.txt
for f in $O/s8b_*.json; do echo "$f $(python -c "
import json
try:
    d=json.load(open('$f')); print('%.0f clips/s %.3f ms gemm %.4f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms']))
except Exception as e: print('ERR', e)
")"; done
head -20 $O/s8b_nccl_info.txt | cut -c1-200
