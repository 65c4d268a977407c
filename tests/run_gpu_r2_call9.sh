#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -s -k "fp32" > $O/c9_fp32.log 2>&1; echo "pytest exit $?" >> $O/c9_fp32.log
grep -E "fp32 tier|cfg1 logits|passed|failed|Error" $O/c9_fp32.log | head -20
