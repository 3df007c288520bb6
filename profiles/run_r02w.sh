#!/bin/bash
# Round 2, call w: ReLU bit masks in the small-layer chains: full suite, A/B, launch list.
OUT=gpurun_out/r02w; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout=200 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
bash profiles/run_ab.sh r02w "CC_SMALL_CHAIN=0" "CC_PRECISION=bf16"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-loss-check > $OUT/ncu1.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02w/launches_bench.csv')) if len(r)>14 and r[0].isdigit()]
names=[r[4] for r in rows]; t=[float(r[14])/1e3 for r in rows]
adams=[i for i,n in enumerate(names) if 'adam_kernel' in n]
a,b=adams[3],adams[4]
print('chain kernels', [round(t[i],1) for i in range(a+1,b+1) if 'chain_tc' in names[i]], 'sum of step', round(sum(t[a+1:b+1]),1))
PY
