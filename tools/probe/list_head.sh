#!/bin/bash
# Dev probe (GPU box): per-kernel durations of one scoring call of model type $1 (ncu launch list, cold cache).
mt=$1; n=${2:-4096}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/list_$mt.csv python tools/probe/run_head.py $mt $n > /dev/null 2>&1
python - "$mt" <<'PY'
import csv, sys, collections
mt = sys.argv[1]
rows = [l for l in open(f"gpurun_out/list_{mt}.csv") if l.startswith('"')]
ds = [d for d in csv.DictReader(rows) if d.get("Metric Name") == "gpu__time_duration.sum"]
half = ds[len(ds) // 2:]                 # run_head.py scores twice: the second call
tot = 0.0
for d in half:
    v = float(d["Metric Value"].replace(",", "")) / 1e3
    tot += v
    print(f"  {d['Kernel Name'][:70]:70s} {v:9.1f} us  grid {d['Grid Size']} block {d['Block Size']}")
print(f"{mt}: {len(half)} launches, {tot:.1f} us")
PY
