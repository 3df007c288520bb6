#!/bin/bash
# Round 2, call v: launch list of the step with the small-layer chains.
OUT=gpurun_out/r02v; mkdir -p $OUT
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02v/launches_bench.csv')) if len(r)>14 and r[0].isdigit()]
names=[r[4] for r in rows]; t=[float(r[14])/1e3 for r in rows]
adams=[i for i,n in enumerate(names) if 'adam_kernel' in n]
a,b=adams[3],adams[4]
for i in range(a+1,b+1):
    print(f"{t[i]:8.1f} us grid {rows[i][8]:>14} {names[i].split('(')[0][:70]}")
print('sum', sum(t[a+1:b+1]))
PY
