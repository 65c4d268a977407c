#!/bin/bash
# re-validation after the last library change: smoke(), full GPU suite with the shipped defaults, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/f2_smoke.log 2>&1; echo "exit $?" >> $O/f2_smoke.log
python -m pytest tests -m gpu -q -rA --timeout=1500 > $O/f2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/f2_pytest_gpu.log
timeout 900 python bench.py > $O/f2_bench_default.json 2> $O/f2_bench_default.err
tail -2 $O/f2_smoke.log; tail -3 $O/f2_pytest_gpu.log; head -c 300 $O/f2_bench_default.json
