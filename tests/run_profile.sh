#!/bin/bash
# launch list of one train step + one full ncu capture of the top kernels (1 GPU). Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tests/profile_step.py > gpurun_out/profile_step.log 2>&1
python - <<'PY'
import csv, collections, re
rows = []
with open("gpurun_out/launches.csv") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", name)[:90]
    agg[name][0] += 1; agg[name][1] += us; tot += us
with open("gpurun_out/launch_summary.txt", "w") as o:
    o.write(f"one train step, B=64, cfg2: {sum(a[0] for a in agg.values())} launches, {tot/1000:.3f} ms summed kernel time (ncu, serialised, cold cache)\n")
    o.write(f"{'kernel':92s} {'n':>5s} {'total_us':>10s} {'share':>7s}\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.write(f"{k:92s} {n:5d} {us:10.1f} {us/tot*100:6.2f}%\n")
print(open("gpurun_out/launch_summary.txt").read())
PY
