#!/bin/bash
# final validation of the round: smoke(), full GPU suite with the shipped defaults, default bench line, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; echo "exit $?" >> $O/f_smoke.log
python -m pytest tests -m gpu -q -rA --timeout=1500 > $O/f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/f_pytest_gpu.log
bash tests/run_profile.sh > /dev/null 2>&1; cp $O/launch_summary.txt $O/f_launches_final.txt
timeout 900 python bench.py > $O/f_bench_default.json 2> $O/f_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/f_bench_reference.json 2> $O/f_bench_reference.err
tail -2 $O/f_smoke.log; tail -3 $O/f_pytest_gpu.log; head -1 $O/f_launches_final.txt; head -c 300 $O/f_bench_default.json; echo; head -c 200 $O/f_bench_reference.json
