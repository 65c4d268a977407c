#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="--steps 20 --warmup 5 --stock 0"
runN() { n=$1; port=$2; shift 2; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $B "$@"; }
timeout 300 python bench.py --gpus 1 $B > $O/s8_cfg2_n1.json 2> $O/s8_cfg2_n1.err
runN 4 29611 > $O/s8_cfg2_n4.json 2> $O/s8_cfg2_n4.err
runN 8 29612 > $O/s8_cfg2_n8.json 2> $O/s8_cfg2_n8.err
timeout 300 python bench.py --gpus 1 --config cfg3 $B > $O/s8_cfg3_n1.json 2> $O/s8_cfg3_n1.err
runN 8 29613 --config cfg3 > $O/s8_cfg3_n8.json 2> $O/s8_cfg3_n8.err
timeout 300 python bench.py --gpus 1 --config cfg5 $B > $O/s8_cfg5_n1.json 2> $O/s8_cfg5_n1.err
runN 2 29614 --config cfg5 > $O/s8_cfg5_n2.json 2> $O/s8_cfg5_n2.err
for f in $O/s8_*.json; do echo "$f $(python -c "
import json
try:
    d=json.load(open('$f')); print('%.0f clips/s %.3f ms gemm %.4f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms']))
except Exception as e: print('ERR', e)
")"; done
