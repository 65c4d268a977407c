#!/bin/bash
# 2-GPU call: DDP gradient equivalence test + scaling lines (cfg2 at N=1,2; cfg3 at N=1,2) with and without SM reservation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/ddp_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ddp.py -m gpu -q -rA -s > $O/ddp_pytest.log 2>&1; echo "pytest exit $?" >> $O/ddp_pytest.log
B="--steps 20 --warmup 5 --stock 0"
run2() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 $B "${@:2}"; }
timeout 300 python bench.py --gpus 1 $B > $O/ddp_bench_cfg2_n1.json 2> $O/ddp_bench_cfg2_n1.err
run2 29511 > $O/ddp_bench_cfg2_n2.json 2> $O/ddp_bench_cfg2_n2.err
PASST_DDP_RESERVE=0 NCCL_MAX_CTAS=32 run2 29512 > $O/ddp_bench_cfg2_n2_noreserve.json 2> $O/ddp_bench_cfg2_n2_noreserve.err
PASST_DDP_RESERVE=8 NCCL_MAX_CTAS=8 run2 29513 > $O/ddp_bench_cfg2_n2_reserve8.json 2> $O/ddp_bench_cfg2_n2_reserve8.err
timeout 300 python bench.py --gpus 1 --config cfg3 $B > $O/ddp_bench_cfg3_n1.json 2> $O/ddp_bench_cfg3_n1.err
run2 29514 --config cfg3 > $O/ddp_bench_cfg3_n2.json 2> $O/ddp_bench_cfg3_n2.err
timeout 300 python bench.py --gpus 1 --config cfg5 $B > $O/ddp_bench_cfg5_n1.json 2> $O/ddp_bench_cfg5_n1.err
run2 29515 --config cfg5 > $O/ddp_bench_cfg5_n2.json 2> $O/ddp_bench_cfg5_n2.err
tail -4 $O/ddp_pytest.log
for f in $O/ddp_bench_*.json; do echo "$f $(python -c "
import json
try:
    d=json.load(open('$f')); print('%.0f clips/s %.3f ms gemm %.4f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms']))
except Exception as e: print('ERR', e)
")"; done
